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


if __name__ == "__main__":
  print(build())
