#!/usr/bin/env python
"""Where does the host-buffer (e2e) path spend its time?  lookup alone, insert alone, both overlapped."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommenders_addons_b200 import dynamic_embedding as de  # noqa: E402
import bench as B  # noqa: E402

dev = torch.device("cuda", 0)
B.bind_to_gpu_numa(0)
dim, res, Bn = 64, 20_000_000, 1 << 20
gen = torch.Generator(device=dev).manual_seed(0)
t = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * dim, init_size=2 * res)
for b in range(0, res, 1 << 20):
  r = torch.arange(b, b + (1 << 20), dtype=torch.int64, device=dev)
  t.insert(B.rank_to_key_torch(r), torch.randn(r.numel(), dim, device=dev, generator=gen) * 0.01)
ks = [B.rank_to_key_torch(torch.randperm(res, device=dev, generator=gen)[:Bn]).cpu().pin_memory() for _ in range(4)]
hv = torch.randn(Bn, dim).pin_memory()
hd = torch.zeros(dim).pin_memory()
ho = [torch.empty(Bn, dim).pin_memory() for _ in range(2)]
torch.cuda.synchronize()


def timed(fn, reps=6):
  fn(0)
  t.host_sync()
  t0 = time.perf_counter()
  for i in range(reps):
    fn(i)
    t.host_sync()
  return (time.perf_counter() - t0) / reps * 1e3


def enqueue_only(fn, reps=6):
  t.host_sync()
  t0 = time.perf_counter()
  for i in range(reps):
    fn(i)
  dt = (time.perf_counter() - t0) / reps * 1e3
  t.host_sync()
  return dt


look = lambda i: t.lookup_host_async(ks[i % 4], hd, ho[i % 2])
ins = lambda i: t.insert_host_async(ks[i % 4], hv)
both = lambda i: (look(i + 1), ins(i))
both_rev = lambda i: (ins(i), look(i + 1))
print(json.dumps({"chunk_mb": os.environ.get("DET_HOST_CHUNK_MB", "8"), "lookup_ms": timed(look), "insert_ms": timed(ins),
                  "both_ms": timed(both), "both_insert_first_ms": timed(both_rev),
                  "enqueue_lookup_ms": enqueue_only(look), "enqueue_insert_ms": enqueue_only(ins)}))
