"""Round-2 measurement of the capacity management (csrc/evict.cu) on a B200 -- run under gpurun AFTER
scripts/gpu_validate_evict.sh is green:
  python scripts/evict_microbench.py [--capacity 100000000] [--dim 64] [--batch 1048576] [--new-frac 0.1]
Prints JSON lines: (1) det_insert vs det_insert_scored on a table below its limit (cost of the score write),
(2) explicit eviction events of k keys: total time and per-phase launch counts, (3) steady state at the limit:
average step time with the events amortised, events per 100 steps, fraction of steps that took the sync path."""
import argparse
import json
import time

import numpy as np
import torch

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommenders_addons_b200 import dynamic_embedding as de  # noqa: E402


def ev_time(fn, reps=5):
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
  ap.add_argument("--capacity", type=int, default=100_000_000)
  ap.add_argument("--dim", type=int, default=64)
  ap.add_argument("--batch", type=int, default=1 << 20)
  ap.add_argument("--new-frac", type=float, default=0.1)
  ap.add_argument("--steps", type=int, default=200)
  ap.add_argument("--fill-frac", type=float, default=0.87, help="resident keys / capacity before the measurements (the soft "
                  "limit of a table with an eviction strategy is 0.875: just below it, so that the steady state evicts)")
  ap.add_argument("--only-events", action="store_true", help="stop after the explicit eviction events (ncu launch lists)")
  a = ap.parse_args()
  dev = torch.device("cuda", 0)
  g = torch.Generator(device=dev)
  g.manual_seed(42)

  def table(strategy, name):
    cfg = de.HkvHashTableConfig(init_capacity=a.capacity, max_capacity=a.capacity, evict_strategy=strategy)
    return de.HkvHashTable(torch.int64, torch.float32, torch.zeros(a.dim), name=name, config=cfg, device=dev)

  vals = torch.randn(a.batch, a.dim, device=dev)
  # (1) score write on the fast path
  plain = de.CuckooHashTable(torch.int64, torch.float32, torch.zeros(a.dim), init_size=a.capacity, device=dev,
                             max_capacity=a.capacity)
  lru = table(de.HkvEvictStrategy.LRU, "mb_lru")
  keys = torch.randint(0, 1 << 62, (a.batch,), device=dev, generator=g)
  for t, name in ((plain, "det_insert"), (lru, "det_insert_scored[LRU]")):
    t.insert(keys, vals)
    ms = ev_time(lambda: t.insert(keys, vals))
    print(json.dumps({"what": "insert below the limit", "path": name, "ms": ms, "keys_per_s": a.batch / ms * 1e3}))
  del plain
  # (2) explicit events
  fill = int(a.capacity * a.fill_frac)
  done = a.batch
  while done < fill:
    k = torch.randint(0, 1 << 62, (a.batch,), device=dev, generator=g)
    lru.insert(k, vals)
    done += a.batch
  for k_ev in (a.batch // 8, a.batch, 4 * a.batch):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = lru.evict(k_ev)
    torch.cuda.synchronize()
    print(json.dumps({"what": "explicit eviction event", "k": k_ev, "evicted": got, "ms": (time.perf_counter() - t0) * 1e3,
                      "resident": int(lru.size())}))
  if a.only_events:
    return
  # (3) steady state at the limit
  n_new = int(a.batch * a.new_frac)
  resident = lru.export()[0]
  torch.cuda.synchronize()
  st0 = lru.stats()
  t0 = time.perf_counter()
  for step in range(a.steps):
    new = torch.randint(0, 1 << 62, (n_new,), device=dev, generator=g)
    old = resident[torch.randint(0, resident.numel(), (a.batch - n_new,), device=dev, generator=g)]
    ks = torch.unique(torch.cat([new, old]))
    lru.insert(ks, vals[:ks.numel()])
  torch.cuda.synchronize()
  dt = (time.perf_counter() - t0) / a.steps * 1e3
  st1 = lru.stats()
  print(json.dumps({"what": "steady state at the limit", "ms_per_step": dt, "steps": a.steps,
                    "events": st1["evict_events"] - st0["evict_events"],
                    "evicted_keys": st1["evicted_keys"] - st0["evicted_keys"], "resident": st1["size"],
                    "used_slots": st1["used_slots"], "capacity": st1["capacity"]}))


if __name__ == "__main__":
  main()
