"""Round-2 probe: what does ONE random read of 32 / 64 / 128 / 256 contiguous bytes cost in DRAM traffic on this GPU?
ncu showed find_kernel_tma reading ~137 B of DRAM per 64 B bucket probe (profiles/r01_find_kernel_tma_dim64.csv: 408.5 MB
per 1,048,576 keys = 8 B key + 256 B row + ~137 B probe); if a random 64 B read really fetches 128 B, 16-slot / 128 B
buckets cost nothing extra per probe (follow-up 1) and 4-slot / 32 B buckets would save nothing.
Run under ncu, one launch per row size (the torch gather kernel is the measured kernel):
  ncu --metrics dram__bytes_read.sum,dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum \\
      --clock-control none -k "regex:[iI]ndex|gather" --csv --log-file gpurun_out/granularity.csv python scripts/probe_granularity.py
Prints the row sizes in launch order; divide each launch's dram__bytes_read by N (and subtract the 8 B index read)."""
import json

import torch

N = 1 << 22
ROWS = 1 << 26  # 64M rows: 2..16 GB tables, far beyond L2


def set_l2_fetch(nbytes):
  """cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity = 0x05, nbytes) on the primary context (what DET_L2_FETCH does
  at table creation); L2_FETCH=32|64|128 in the environment selects it, unset = device default"""
  import ctypes
  import glob
  import os
  cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")) + \
      ["libcudart.so.12", "libcudart.so"]
  for c in cands:
    try:
      rt = ctypes.CDLL(c)
      break
    except OSError:
      continue
  else:
    raise RuntimeError("libcudart not found")
  torch.zeros(1, device="cuda")  # create the primary context
  rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(nbytes))
  got = ctypes.c_size_t(0)
  rt.cudaDeviceGetLimit(ctypes.byref(got), 5)
  return rc, got.value


def main():
  import os
  if os.environ.get("L2_FETCH"):
    print(json.dumps({"l2_fetch_request": int(os.environ["L2_FETCH"]), "rc_and_value": set_l2_fetch(int(os.environ["L2_FETCH"]))}))
  dev = torch.device("cuda", 0)
  g = torch.Generator(device=dev).manual_seed(0)
  order = []
  for row_bytes in (32, 64, 128, 256):
    words = row_bytes // 8
    table = torch.zeros((ROWS, words), dtype=torch.int64, device=dev)
    idx = torch.randint(0, ROWS, (N,), device=dev, generator=g)
    torch.cuda.synchronize()
    for _ in range(2):
      out = table.index_select(0, idx)
    torch.cuda.synchronize()
    order.append({"row_bytes": row_bytes, "reads": N, "launches": 2, "out_bytes": int(out.numel() * 8)})
    del table, out
    torch.cuda.empty_cache()
  # second question: is part of the per-probe DRAM traffic PAGE-TABLE WALKS?  Random 64 B reads over growing footprints:
  # a cost per read that rises with the footprint is TLB-miss traffic, not fetch granularity.  The engine touches two
  # random pages per key (key plane + value plane of a 53 GB table); if the walks show up here, a bucket-major layout
  # (a bucket's keys next to its rows, one page per key) is the follow-up, not a different bucket width.
  for gib in (0.5, 4, 32, 96):
    rows = int(gib * (1 << 30)) // 64
    free = torch.cuda.mem_get_info()[0]
    if rows * 64 + (2 << 30) > free:
      continue
    table = torch.zeros((rows, 8), dtype=torch.int64, device=dev)
    idx = torch.randint(0, rows, (N,), device=dev, generator=g)
    torch.cuda.synchronize()
    for _ in range(2):
      out = table.index_select(0, idx)
    torch.cuda.synchronize()
    order.append({"row_bytes": 64, "footprint_gib": gib, "reads": N, "launches": 2, "out_bytes": int(out.numel() * 8)})
    del table, out
    torch.cuda.empty_cache()
  print(json.dumps(order))


if __name__ == "__main__":
  main()
