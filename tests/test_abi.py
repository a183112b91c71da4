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
  assert lib.det_abi_version() == 1
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
