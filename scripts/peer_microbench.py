#!/usr/bin/env python
"""Single-process, 2-GPU probe of the one-sided path: shard 0 on cuda:0, shard 1 on cuda:1 (plain peer access),
kernels launched on cuda:0.  Times det_peer_find / det_peer_insert for keys owned by the local shard, the remote
shard, and a mix, for a TLB-friendly and a BASELINE-sized table."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recommenders_addons_b200 import dynamic_embedding as de  # noqa: E402
import bench as B  # noqa: E402


def t_ms(fn, reps=6):
  ts = []
  for _ in range(reps + 2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  return float(np.median(ts[2:]))


def main():
  dim, Bn = 64, 1 << 20
  torch.cuda.set_device(0)
  d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
  # raw peer copy bandwidth for reference
  x0 = torch.empty(1 << 28, dtype=torch.uint8, device=d0)
  x1 = torch.empty(1 << 28, dtype=torch.uint8, device=d1)
  ms = t_ms(lambda: x0.copy_(x1))
  print(json.dumps({"op": "torch peer copy d1->d0 256MiB", "ms": ms, "GBs": (1 << 28) / ms / 1e6}), flush=True)
  del x0, x1
  for res in (1 << 20, 50_000_000):
    shards = [de.Variable(dim=dim, init_size=2 * res, initializer=0.0, name="pm-%d-%d" % (res, i), devices=["cuda:%d" % i])
              for i in range(2)]
    pv = de.PeerShardedVariable(fake_shards=shards)
    gen = torch.Generator(device=d0).manual_seed(1)
    # fill both shards through the one-sided path (keys of all owners), 1M at a time
    own_keys = [[], []]
    for b in range(0, 2 * res, 1 << 20):
      r = torch.arange(b, b + (1 << 20), dtype=torch.int64, device=d0)
      k = B.rank_to_key_torch(r)
      pv.upsert(k, torch.randn(k.numel(), dim, device=d0, generator=gen) * 0.01)
      if b < (8 << 20):
        o = (k & 0x7fffffff) % 2
        own_keys[0].append(k[o == 0])
        own_keys[1].append(k[o == 1])
    torch.cuda.synchronize()
    k_local = torch.cat(own_keys[0])[:Bn].contiguous()
    k_remote = torch.cat(own_keys[1])[:Bn].contiguous()
    # spread over the whole table, not just the first chunks
    allr = torch.randperm(2 * res, device=d0, generator=gen)[:4 * Bn]
    ka = B.rank_to_key_torch(allr)
    oa = (ka & 0x7fffffff) % 2
    k_local, k_remote = ka[oa == 0][:Bn].contiguous(), ka[oa == 1][:Bn].contiguous()
    k_mix = ka[:Bn].contiguous()
    vals = torch.randn(Bn, dim, device=d0, generator=gen) * 0.01
    sizes = [int(s.size()) for s in shards]
    for name, k in (("local", k_local), ("remote", k_remote), ("mix", k_mix)):
      n = k.numel()
      ms_f = t_ms(lambda: pv.lookup(k))
      ms_i = t_ms(lambda: pv.upsert(k, vals[:n]))
      hit = float(pv.lookup(k, return_exists=True)[1].float().mean())
      print(json.dumps({"resident_per_shard": res, "sizes": sizes, "keys": name, "n": n, "hit": hit,
                        "find_ms": round(ms_f, 4), "find_Mkeys_s": round(n / ms_f / 1e3, 1),
                        "find_remote_GBs": round(n * (64 + dim * 4) / ms_f / 1e6, 1),
                        "insert_ms": round(ms_i, 4), "insert_Mkeys_s": round(n / ms_i / 1e3, 1)}), flush=True)
    # the plain single-table kernels on the same local shard, for reference
    t0 = shards[0].tables[0]
    ms_f = t_ms(lambda: t0.lookup(k_local))
    print(json.dumps({"resident_per_shard": res, "keys": "local via det_find", "find_ms": round(ms_f, 4),
                      "find_Mkeys_s": round(k_local.numel() / ms_f / 1e3, 1)}), flush=True)
    pv.close()
    for s in shards:
      s.tables[0].close()
    del shards, pv
    torch.cuda.empty_cache()


if __name__ == "__main__":
  main()
