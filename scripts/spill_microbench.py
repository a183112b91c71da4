"""Round-2 measurement of `max_hbm_for_vectors` on a B200 -- run under gpurun after
scripts/gpu_validate_evict.sh is green:
  python scripts/spill_microbench.py [--keys 20000000] [--dim 64] [--batch 1048576]
For HBM fractions 1.0 / 0.75 / 0.5 / 0.0 of the value plane: find and insert (resident keys) time per batch, keys/s,
and the implied PCIe traffic (rows on the host side x row bytes).  Expectation: the host fraction of a uniform batch
moves at the PCIe gather rate (<= ~55 GB/s per direction), the HBM fraction at the usual rate."""
import argparse
import json

import numpy as np
import torch

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommenders_addons_b200 import dynamic_embedding as de  # noqa: E402


def ev_time(fn, reps=7):
  out = []
  for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    out.append(a.elapsed_time(b))
  return float(np.median(out))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--keys", type=int, default=20_000_000)
  ap.add_argument("--dim", type=int, default=64)
  ap.add_argument("--batch", type=int, default=1 << 20)
  a = ap.parse_args()
  dev = torch.device("cuda", 0)
  g = torch.Generator(device=dev)
  g.manual_seed(42)
  cap = int(a.keys / 0.5)
  row = a.dim * 4
  for frac in (1.0, 0.75, 0.5, 0.0):
    budget = None if frac == 1.0 else max(int(cap * row * frac), 1)
    t = de.CuckooHashTable(torch.int64, torch.float32, torch.zeros(a.dim), init_size=cap, device=dev, max_capacity=cap,
                           max_hbm_for_values=budget)
    st = t.stats()
    keys = torch.randperm(a.keys, device=dev, generator=g) * 2654435761 + 17
    for lo in range(0, a.keys, a.batch):
      k = keys[lo:lo + a.batch]
      t.insert(k, torch.randn(k.numel(), a.dim, device=dev))
    q = keys[torch.randint(0, a.keys, (a.batch,), device=dev, generator=g)].unique()
    vals = torch.randn(q.numel(), a.dim, device=dev)
    t.lookup(q)
    f_ms = ev_time(lambda: t.lookup(q))
    i_ms = ev_time(lambda: t.insert(q, vals))
    host_frac = st["host_bytes"] / max(st["host_bytes"] + (st["capacity"] + 2) * row - st["host_bytes"], 1)
    print(json.dumps({"hbm_frac": frac, "hbm_bytes": st["hbm_bytes"], "host_bytes": st["host_bytes"], "batch": q.numel(),
                      "find_ms": f_ms, "find_keys_per_s": q.numel() / f_ms * 1e3, "insert_ms": i_ms,
                      "insert_keys_per_s": q.numel() / i_ms * 1e3,
                      "pcie_gb_per_s_find": host_frac * q.numel() * row / f_ms / 1e6,
                      "pcie_gb_per_s_insert": host_frac * q.numel() * row / i_ms / 1e6}), flush=True)
    del t
    torch.cuda.empty_cache()


if __name__ == "__main__":
  main()
