"""First thing to run on a B200 for the sharded-table-with-eviction path: it was written after the round's
GPU budget was spent and has only run on the SIMT emulator (tests/test_host_peer_emu.py).  Standalone on purpose -- not
collected by pytest -- so an unvalidated path cannot take the validated GPU suite down with it.

  single GPU:   python scripts/gpu_sharded_evict.py
  N GPUs:       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
                  --master-port 29571 scripts/gpu_sharded_evict.py

Every rank trains 30 HOT ids every step plus fresh cold ids through det_peer_xchg_apply_adagrad on fixed-capacity shards
with an LFU strategy; the shards must stay under their hard bound, count eviction events, keep every hot id with the
parameters of a sequential Adagrad model (bit for bit) and raise no error flag.  Prints one JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
  import torch.distributed as dist
  from recommenders_addons_b200 import dynamic_embedding as de
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
  dev = torch.device("cuda", torch.cuda.current_device())
  if world > 1:
    dist.init_process_group("nccl")
  dim, slots, cap, steps = 64, 1 << 14, 1024, 120
  f32 = np.float32
  rng = np.random.default_rng(4)                       # the SAME schedule on every rank
  hot = rng.choice(1 << 40, size=30, replace=False).astype(np.int64)
  sched = []
  for t in range(steps):
    per = []
    for r in range(world):
      cold = (rng.choice(1 << 40, size=cap - 30, replace=False).astype(np.int64) | (1 << 41)) + (t * world + r) * (1 << 42)
      per.append((np.concatenate([hot, cold]), rng.normal(0, 1e-2, (cap, dim)).astype(f32)))
    sched.append(per)
  if world > 1:
    pv = de.PeerShardedVariable.create(dim, slots, initializer=0.05, num_slot_planes=1, name="gpu-evict",
                                       evict_strategy=de.HkvEvictStrategy.LFU)
    pv.attach_exchange(cap, insert="push")
  else:
    cfg = de.HkvHashTableConfig(init_capacity=slots, max_capacity=slots, evict_strategy=de.HkvEvictStrategy.LFU)
    var = de.Variable(dim=dim, init_size=slots, initializer=0.05, num_slot_planes=1, name="gpu-evict",
                      kv_creator=de.HkvHashTableCreator(config=cfg))
    pv = de.PeerShardedVariable(fake_shards=[var])
    nbytes = int(pv._lib.det_peer_xchg_bytes(1, cap, dim * 4))
    raw = torch.zeros(nbytes + 256, dtype=torch.uint8, device=dev)
    off = (-raw.data_ptr()) % 256
    pv.attach_exchange(cap, mailbox_ptrs=[raw.data_ptr() + off], keepalive=raw, insert="push")
  opt = de.FusedAdagrad(0.1, 0.1)
  torch.cuda.synchronize()
  t0 = time.time()
  for t in range(steps):
    k, g = sched[t][rank]
    pv.apply_gradients(opt, torch.from_numpy(k).to(dev), torch.from_numpy(g).to(dev))
  torch.cuda.synchronize()
  dt = time.time() - t0
  # sequential model of the hot ids (summed over ranks in source order)
  ip = np.full(dim, 0.05, f32)
  par = {int(kk): ip.copy() for kk in hot}
  acc = {int(kk): np.full(dim, 0.1, f32) for kk in hot}
  for t in range(steps):
    gs = {}
    for r in range(world):
      for kk, row in zip(sched[t][r][0][:30].tolist(), sched[t][r][1][:30]):
        gs[kk] = row.copy() if kk not in gs else (gs[kk] + row).astype(f32)
    for kk, gg in gs.items():
      acc[kk] = (acc[kk] + gg * gg).astype(f32)
      par[kk] = (par[kk] - (f32(0.1) * gg) / np.sqrt(acc[kk])).astype(f32)
  # upsert through the owners (det_peer_xchg_insert on evicting shards): 8 rank-private keys per rank, read back by all
  mine_k = torch.arange(8, dtype=torch.int64, device=dev) + 1000 * (rank + 1)
  pv.upsert(mine_k, torch.full((8, dim), float(rank + 1), device=dev))
  allk = torch.cat([torch.arange(8, dtype=torch.int64, device=dev) + 1000 * (r + 1) for r in range(world)])
  up = pv.lookup(allk)
  upsert_ok = bool((up[:, 0] == (allk // 1000).to(torch.float32)).all())
  rows, ex = pv.lookup(torch.from_numpy(hot).to(dev), return_exists=True)
  rows, ex = rows.cpu().numpy(), ex.cpu().numpy()
  exp = np.stack([par[int(kk)] for kk in hot])
  local = pv.local.tables[0]
  st = local.stats()
  lk, _ = local.export()
  mine = bool((de.default_partition_fn(lk, world, True) == rank).all())
  ok = bool(ex.all()) and bool(np.array_equal(rows, exp)) and st["error_flags"] == 0 and st["evict_events"] > 0 and \
      int(local.size()) <= int(slots * 0.95) and mine and upsert_ok
  res = {"rank": rank, "world": world, "ok": ok, "hot_found": int(ex.sum()), "upsert_ok": upsert_ok, "hot_rows_bit_exact": bool(np.array_equal(rows, exp)),
         "size": int(local.size()), "slots": slots, "evict_events": int(st["evict_events"]), "error_flags": int(st["error_flags"]),
         "ms_per_step": 1e3 * dt / steps}
  print(json.dumps(res), flush=True)
  pv.close()
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  return 0 if ok else 1


if __name__ == "__main__":
  sys.exit(main())
