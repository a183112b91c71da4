"""Test backend: the Python mirror (recommenders_addons_b200.dynamic_embedding) running over the EMULATED libdetable
(tests/emu/build_emu.py: table.cu + fused.cu + evict.cu compiled by g++ against the SIMT emulator) with CPU tensors as
"device" memory.  It lets the Python glue -- table / Variable / optimizer / restrict-policy code -- of GPU test bodies be
executed without a GPU.  Installed by monkeypatching, only inside tests; the product has no CPU path."""
import contextlib
import ctypes

from recommenders_addons_b200 import _lib
from recommenders_addons_b200.dynamic_embedding import ops, optimizer, sharded, table, variable
from tests.emu import build_emu

_EMU = None


def emu_cdll():
  global _EMU
  if _EMU is None:
    l = ctypes.CDLL(build_emu.build_lib())
    for name, (res, args) in _lib.SIGNATURES.items():
      fn = getattr(l, name, None)      # host-buffer and peer entry points are not part of the emulated build
      if fn is not None:
        fn.restype, fn.argtypes = res, args
    _EMU = l
  return _EMU


@contextlib.contextmanager
def installed():
  saved = (_lib._LIB, table._DEVICE_TYPES, table._stream_ptr, variable._stream_ptr, optimizer._stream_ptr,
           ops._stream_ptr)
  saved_sp = sharded.PeerShardedVariable._sp
  sharded.PeerShardedVariable._sp = lambda self: None
  _lib._LIB = emu_cdll()
  table._DEVICE_TYPES = ("cuda", "cpu")
  no_stream = lambda device: None  # noqa: E731  (the emulated runtime is synchronous)
  table._stream_ptr = variable._stream_ptr = optimizer._stream_ptr = ops._stream_ptr = no_stream
  variable._reset_variables()
  try:
    yield
  finally:
    variable._reset_variables()
    sharded.PeerShardedVariable._sp = saved_sp
    (_lib._LIB, table._DEVICE_TYPES, table._stream_ptr, variable._stream_ptr, optimizer._stream_ptr,
     ops._stream_ptr) = saved
