#!/usr/bin/env python
"""NVLink store probe (scripts/nvlink_probe.cu): GB/s of every way of moving 256 B rows to a peer GPU, one direction and
both directions at once.  Run under `gpurun --gpus 2`; prints JSON lines (kept in profiles/r02_nvlink_probe.jsonl)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "recommenders_addons_b200", "lib", "libdetprobe.so"))
lib.probe_run.restype = ctypes.c_int
lib.probe_run.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                          ctypes.POINTER(ctypes.c_double)]
NAMES = {0: "cudaMemcpyPeerAsync", 1: "st.v4 stream contiguous", 2: "st.v4 rows -> RANDOM remote rows (from registers)",
         3: "local RANDOM rows -> remote contiguous (st.v4)", 4: "local contiguous -> remote RANDOM rows (st.v4)",
         5: "TMA bulk store, contiguous chunks", 6: "TMA bulk store, single rows to RANDOM remote rows",
         7: "remote RANDOM row READ -> local (pull)"}


def run(mode, chunk=4096, ctas=4, bidir=0, nbytes=1 << 30, reps=10):
  out = (ctypes.c_double * 2)()
  rc = lib.probe_run(mode, nbytes, chunk, ctas, bidir, reps, out)
  rec = {"mode": mode, "what": NAMES[mode], "chunk": chunk if mode in (5, 6) else None, "ctas_per_sm": ctas, "bidir": bool(bidir),
         "GBs_0to1": round(out[0], 1), "GBs_1to0": round(out[1], 1), "rc": rc}
  print(json.dumps(rec), flush=True)


if __name__ == "__main__":
  for bidir in (0, 1):
    run(0, bidir=bidir)
    for ctas in (2, 4, 8):
      run(1, ctas=ctas, bidir=bidir)
    for mode in (2, 3, 4, 7):
      for ctas in (4, 8):
        run(mode, ctas=ctas, bidir=bidir)
    for chunk in (256, 1024, 4096, 16384, 65536):
      for ctas in (2, 4) if chunk >= 16384 else (4, 8):
        run(5, chunk=chunk, ctas=ctas, bidir=bidir)
    for chunk in (4096, 16384):
      run(6, chunk=chunk, ctas=4, bidir=bidir)
