"""Builds tests/emu/_build/libevict_emu.so: the device code of csrc/{common.cuh,evict_kernels.cuh} compiled by g++
under the SIMT emulator (cuda_emu.h).  Test infrastructure only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libevict_emu.so")
SRC = os.path.join(HERE, "evict_emu.cc")
DEPS = [SRC, os.path.join(HERE, "cuda_emu.h"),
        os.path.join(ROOT, "recommenders_addons_b200", "csrc", "common.cuh"),
        os.path.join(ROOT, "recommenders_addons_b200", "csrc", "evict_kernels.cuh"),
        os.path.join(ROOT, "include", "detable.h")]


def build():
  if os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
    return OUT
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  tmp = OUT + ".tmp.%d" % os.getpid()
  cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DDET_EMU=1", "-Wall", "-Wno-unused-function",
         "-Wno-unknown-pragmas", "-o", tmp, SRC]
  subprocess.run(cmd, check=True)
  os.replace(tmp, OUT)
  return OUT


CSRC = os.path.join(ROOT, "recommenders_addons_b200", "csrc")
LIB_OUT = os.path.join(HERE, "_build", "libdetable_emu.so")
LIB_SRCS = [os.path.join(CSRC, f) for f in ("table.cu", "fused.cu", "evict.cu", "host_api.cu", "sharded.cu")]
LIB_DEPS = LIB_SRCS + [os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "cuda_runtime_emu.h"),
                       os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "host.h"),
                       os.path.join(CSRC, "evict_kernels.cuh"), os.path.join(ROOT, "include", "detable.h")]


def _sanitize_flags():
  """DET_EMU_SANITIZE=address,undefined (scripts/emu_sanitize.sh): the emulated library instrumented by
  AddressSanitizer / UBSan -- out-of-bounds and misaligned accesses of the KERNELS are caught on the CPU, because
  "device" memory is malloc'd host memory with red zones.  The process must preload the sanitizer runtimes."""
  what = os.environ.get("DET_EMU_SANITIZE", "")
  return ["-fsanitize=" + what, "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer"] if what else []


def build_lib():
  """ALL of the engine's translation units (host code AND kernels) compiled by g++ against the emulator: the real C ABI
  (det_table_create, det_find, det_insert, det_insert_scored, ...) runs on the CPU"""
  global LIB_OUT
  san = _sanitize_flags()
  if san and not LIB_OUT.endswith("_san.so"):
    LIB_OUT = LIB_OUT[:-3] + "_san.so"
  if os.path.exists(LIB_OUT) and all(os.path.getmtime(LIB_OUT) >= os.path.getmtime(d) for d in LIB_DEPS):
    return LIB_OUT
  os.makedirs(os.path.dirname(LIB_OUT), exist_ok=True)
  tmp = LIB_OUT + ".tmp.%d" % os.getpid()
  cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-DDET_EMU=1", "-Wall", "-Wno-unused-function",
         "-Wno-unknown-pragmas", "-Wno-unused-variable"] + san + ["-include", os.path.join(HERE, "cuda_emu.h"), "-include",
         os.path.join(HERE, "cuda_runtime_emu.h"), "-x", "c++", "-o", tmp] + LIB_SRCS
  subprocess.run(cmd, check=True)
  os.replace(tmp, LIB_OUT)
  return LIB_OUT


if __name__ == "__main__":
  print(build())
  print(build_lib())
