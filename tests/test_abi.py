"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/detable.h
declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "detable.h")


def declared_functions():
  src = open(HEADER).read()
  src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
  return sorted(set(re.findall(r"\b(det_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_expected_surface():
  names = declared_functions()
  for must in ("det_table_create", "det_table_destroy", "det_find", "det_insert", "det_accum", "det_remove",
               "det_clear", "det_size", "det_export", "det_import", "det_lookup_sparse", "det_apply_adagrad",
               "det_apply_adam", "det_unique", "det_partition", "det_find_host", "det_insert_host", "det_save",
               "det_load", "det_last_error"):
    assert must in names


def test_library_exports_every_declared_symbol():
  from recommenders_addons_b200 import _lib
  lib = _lib.lib()
  names = declared_functions()
  assert set(names) == set(_lib.SIGNATURES.keys()), set(names) ^ set(_lib.SIGNATURES.keys())
  for n in names:
    assert getattr(lib, n) is not None
  assert lib.det_abi_version() == _lib.ABI_VERSION == 8
  assert b"sm_100a" in lib.det_build_info()


def test_library_contains_sm100a_code_only():
  from recommenders_addons_b200 import build
  path = build.build()
  cuobjdump = "/usr/local/cuda/bin/cuobjdump"
  if not os.path.exists(cuobjdump):
    pytest.skip("cuobjdump not available")
  out = subprocess.run([cuobjdump, "-lelf", path], capture_output=True, text=True).stdout
  archs = set(re.findall(r"sm_\d+a?", out))
  assert archs == {"sm_100a"}, archs


def test_header_is_plain_c():
  """The boundary must compile as C: no C++ / torch types in the signatures."""
  test_c = "#include \"%s\"\nint main(void){ det_config c; (void)c; return DET_OK; }\n" % HEADER
  p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", "-"], input=test_c,
                     capture_output=True, text=True)
  assert p.returncode == 0, p.stderr


def test_error_convention_without_gpu():
  """Null arguments are rejected with DET_INVALID_ARGUMENT and a message, before any CUDA call."""
  from recommenders_addons_b200 import _lib
  lib = _lib.lib()
  assert lib.det_table_create(None, None) == 1
  assert b"null" in lib.det_last_error()
  cfg = _lib.DetConfig()
  cfg.value_dtype, cfg.dim = 99, 4
  h = ctypes.c_void_p()
  assert lib.det_table_create(ctypes.byref(h), ctypes.byref(cfg)) == 1
  assert b"value_dtype" in lib.det_last_error()
  cfg.value_dtype, cfg.dim = 0, 0
  assert lib.det_table_create(ctypes.byref(h), ctypes.byref(cfg)) == 1
  assert b"dim" in lib.det_last_error()


def test_ctypes_structs_match_the_c_header(tmp_path):
  """sizeof/offsetof of det_config and det_stats as the C compiler sees them == the ctypes mirrors in _lib.py."""
  from recommenders_addons_b200 import _lib
  src = tmp_path / "layout.c"
  src.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "%s"
int main(void) {
  printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n", sizeof(det_config), offsetof(det_config, init_capacity),
         offsetof(det_config, max_load_factor), sizeof(det_stats), offsetof(det_stats, hbm_bytes),
         offsetof(det_stats, rehash_count), offsetof(det_config, max_hbm_for_vectors), offsetof(det_stats, host_bytes));
  return 0;
}
''' % HEADER)
  exe = tmp_path / "layout"
  p = subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], capture_output=True, text=True)
  assert p.returncode == 0, p.stderr
  out = subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()
  got = [int(x) for x in out]
  exp = [ctypes.sizeof(_lib.DetConfig), _lib.DetConfig.init_capacity.offset, _lib.DetConfig.max_load_factor.offset,
         ctypes.sizeof(_lib.DetStats), _lib.DetStats.hbm_bytes.offset, _lib.DetStats.rehash_count.offset,
         _lib.DetConfig.max_hbm_for_vectors.offset, _lib.DetStats.host_bytes.offset]
  assert got == exp, (got, exp)


def test_peer_and_region_entry_points_reject_bad_arguments_without_gpu():
  from recommenders_addons_b200 import _lib
  lib = _lib.lib()
  assert lib.det_peer_handle_bytes() > 64 * 3
  assert lib.det_peer_inbox_bytes(8, 1 << 20, 256) == 256 + 8 * ((1 << 23) + (1 << 28))
  assert lib.det_peer_inbox_bytes(9, 10, 16) == 0          # at most 8 GPUs in one NVSwitch domain
  cfg = _lib.DetConfig()
  cfg.value_dtype, cfg.dim, cfg.init_capacity, cfg.num_slot_planes = 0, 64, 1 << 20, 1
  n = lib.det_table_region_bytes(ctypes.byref(cfg))
  assert n >= (1 << 20) * (8 + 2 * 256) and n % 256 == 0
  h = ctypes.c_void_p()
  assert lib.det_table_create_in_region(ctypes.byref(h), ctypes.byref(cfg), None, 0) == 1
  assert lib.det_peer_group_create_regions(ctypes.byref(h), None, None, 2, 0, 1) == 1
  assert lib.det_peer_route(None, None, None, 0, None) == 1
  assert lib.det_host_sync(None) == 1


def test_scored_entry_points_reject_bad_arguments_without_gpu():
  """the capacity-management entry points (det_insert_scored ...) validate their arguments before touching CUDA"""
  from recommenders_addons_b200 import _lib
  lib = _lib.lib()
  assert lib.det_insert_scored(None, None, None, None, 0, None) == 1
  assert lib.det_accum_scored(None, None, None, None, None, 0, None) == 1
  assert lib.det_find_scores(None, None, 0, None, None) == 1
  assert lib.det_set_global_epoch(None, 3) == 1
  assert lib.det_evict(None, 1, None, None) == 1
  assert b"null table" in lib.det_last_error()
  assert _lib.flags_evict(_lib.EVICT_STRATEGIES["LRU"]) == 1 and _lib.flags_evict(-1) == 0
  assert ctypes.sizeof(_lib.DetStats) == 72
