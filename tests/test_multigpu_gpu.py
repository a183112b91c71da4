"""Real multi-GPU paths (need >= 2 GPUs on the box; run with `gpurun --gpus 2`): one process per GPU,
key-hash sharded table.  Checks (a) the NCCL all-to-all exchange (ShardedVariable) and (b) the one-sided NVLink
peer-memory path (PeerShardedVariable) against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
  import torch
  import torch.distributed as dist
  from oracle import oracle as O
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
  from recommenders_addons_b200 import dynamic_embedding as de
  dev = torch.device("cuda", rank)
  dim = 64
  ok = True
  msg = ""
  try:
    for mode in ("nccl", "peer", "symm", "push", "hybrid"):
      if mode in ("symm", "push", "hybrid"):  # shard inside a torch symmetric-memory region (CUDA VMM): the production path
        sv = de.PeerShardedVariable.create(dim, 1 << 18, initializer=-1.0, name="mg-%s-%d" % (mode, rank))
        var = sv.local
        if mode == "push":          # owner-side exchange: ids to the owner, local probe, rows pushed back; no barriers
          sv.attach_exchange(1 << 16)
        if mode == "hybrid":        # lookups through the owners, one-sided inserts, flag barrier between the phases
          sv.attach_exchange(1 << 16, insert="pull")
      else:
        var = de.Variable(dim=dim, init_size=1 << 18, initializer=-1.0, name="mg-%s-%d" % (mode, rank))
        sv = de.ShardedVariable(var) if mode == "nccl" else de.PeerShardedVariable(var)
      # every rank writes its own disjoint key set (owned by arbitrary ranks) ...
      rng = np.random.default_rng(1234)
      allkeys = rng.choice(np.arange(-10**6, 10**6), 60000, replace=False).astype(np.int64)
      allvals = rng.normal(0, 0.01, (allkeys.shape[0], dim)).astype(np.float32)
      mine = slice(rank, None, world)
      sv.upsert(torch.from_numpy(allkeys[mine]).to(dev), torch.from_numpy(allvals[mine]).to(dev))
      if mode != "nccl":
        sv.phase_barrier()
      torch.cuda.synchronize()
      dist.barrier()
      # ... and reads keys written by EVERY rank
      ot = O.PortTable(dim)
      ot.insert(allkeys, allvals)
      q_keys = np.concatenate([allkeys[::3], rng.integers(2 * 10**6, 3 * 10**6, 4000)]).astype(np.int64)
      np.random.default_rng(rank).shuffle(q_keys)
      got = sv.lookup(torch.from_numpy(q_keys).to(dev))
      exp = ot.find(q_keys, np.full(dim, -1, np.float32))
      if not np.array_equal(got.reshape(-1, dim).cpu().numpy(), exp):
        ok, msg = False, "lookup mismatch in mode %s" % mode
      own = int(var.size())
      owner = O.default_partition_fn(allkeys, world, True)
      if own != int((owner == rank).sum()):
        ok, msg = False, "shard size %d != %d in mode %s" % (own, int((owner == rank).sum()), mode)
      if mode != "nccl":
        sv.phase_barrier()
        torch.cuda.synchronize()
      dist.barrier()
      if mode == "symm":
        # backward path: every rank routes row-gradients for keys of ALL owners; owners combine and step
        var2 = de.PeerShardedVariable.create(dim, 1 << 18, initializer=0.0, num_slot_planes=1, name="mg-opt-%d" % rank)
        var2.attach_inbox(1 << 16)
        opt = de.FusedAdagrad(0.1, 0.1)
        gk = allkeys[:20000]                                   # the SAME keys from both ranks -> summed on the owner
        gg = np.full((gk.shape[0], dim), 0.01 * (rank + 1), np.float32)
        var2.apply_gradients(opt, torch.from_numpy(gk).to(dev), torch.from_numpy(gg).to(dev))
        torch.cuda.synchronize()
        dist.barrier()
        gsum = np.float32(0.01 * sum(range(1, world + 1)))
        acc = np.float32(0.1) + gsum * gsum
        expect = np.float32(0) - (np.float32(0.1) * gsum) / np.sqrt(acc)
        got2 = var2.lookup(torch.from_numpy(gk).to(dev)).cpu().numpy()
        if not np.allclose(got2, expect, rtol=1e-6, atol=1e-8):
          ok, msg = False, "sharded adagrad mismatch: %r vs %r" % (got2[0, 0], expect)
        var2.phase_barrier()
        torch.cuda.synchronize()
        dist.barrier()
        # the caller layer (HvdAllToAllEmbedding mirror): forward over ids of all owners + routed backward
        layer = de.layers.AllToAllEmbedding(dim, 1 << 16, initializer=0.0, name="mg-layer-%d" % rank, num_slot_planes=1)
        layer.train()
        lids = torch.from_numpy(allkeys[:4096].reshape(64, 64)).to(dev)
        out = layer(lids)
        if tuple(out.shape) != (64, 64, dim) or float(out.abs().max()) != 0.0:
          ok, msg = False, "layer forward"
        out.sum().backward()                                   # d/d(row) = number of occurrences = 1 per unique id
        layer.apply_gradients(de.FusedAdagrad(0.1, 0.1), max_unique_per_rank=4096)
        torch.cuda.synchronize()
        dist.barrier()
        gsum = np.float32(world)                                # every rank contributes gradient 1 for the same ids
        expect = np.float32(0) - (np.float32(0.1) * gsum) / np.sqrt(np.float32(0.1) + gsum * gsum)
        got3 = layer.params.lookup(lids.reshape(-1)).cpu().numpy()
        layer.params.phase_barrier()
        torch.cuda.synchronize()
        dist.barrier()
        if not np.allclose(got3, expect, rtol=1e-6, atol=1e-8):
          ok, msg = False, "layer backward: %r vs %r" % (got3[0, 0], expect)
      if mode == "push":
        # several rounds of rewrite -> read with NO barrier in between: the flag words alone must order every owner's
        # reads after the writes of the same round (the horovod_sync_train_test shape: values change every step)
        model = dict(zip(allkeys.tolist(), allvals))
        for rnd in range(1, 4):
          sub = allkeys[mine][rnd::5]
          newv = (allvals[mine][rnd::5] + np.float32(rnd)).astype(np.float32)
          sv.upsert(torch.from_numpy(np.ascontiguousarray(sub)).to(dev), torch.from_numpy(np.ascontiguousarray(newv)).to(dev))
          for r2 in range(world):   # what EVERY rank wrote this round
            k2 = allkeys[slice(r2, None, world)][rnd::5]
            v2 = (allvals[slice(r2, None, world)][rnd::5] + np.float32(rnd)).astype(np.float32)
            model.update(zip(k2.tolist(), v2))
          qk = np.concatenate([allkeys[rnd::7], np.array([5 * 10**6 + rank, -5 * 10**6 - rank], np.int64)])
          got_r, got_e = sv.lookup(torch.from_numpy(qk).to(dev), return_exists=True)
          exp_r = np.stack([model.get(int(k), np.full(dim, -1, np.float32)) for k in qk])
          exp_e = np.array([int(k) in model for k in qk])
          if not (np.array_equal(got_r.cpu().numpy(), exp_r) and np.array_equal(got_e.cpu().numpy(), exp_e)):
            ok, msg = False, "push mode: round %d lookup differs from the model" % rnd
          # per-key default rows (full-size default) for the misses
          fd = torch.from_numpy(np.random.default_rng(rnd).normal(0, 1, (len(qk), dim)).astype(np.float32)).to(dev)
          got_f = sv.lookup(torch.from_numpy(qk).to(dev), default=fd).cpu().numpy()
          exp_f = np.where(exp_e[:, None], exp_r, fd.cpu().numpy())
          if not np.array_equal(got_f, exp_f):
            ok, msg = False, "push mode: full-size defaults in round %d" % rnd
          # zero-copy ring view == copied rows
          v1 = sv.lookup(torch.from_numpy(qk).to(dev), copy=False)
          if not np.array_equal(v1.cpu().numpy(), exp_r):
            ok, msg = False, "push mode: ring view in round %d" % rnd
        torch.cuda.synchronize()
        dist.barrier()
        # the sharded optimizer step through the owners (det_peer_xchg_apply_adagrad): the SAME keys from every rank
        var3 = de.PeerShardedVariable.create(dim, 1 << 18, initializer=0.0, num_slot_planes=1, name="mg-xopt-%d" % rank)
        var3.attach_exchange(1 << 16)
        opt3 = de.FusedAdagrad(0.1, 0.1)
        gk3 = allkeys[:20000]
        gg3 = np.full((gk3.shape[0], dim), 0.01 * (rank + 1), np.float32)
        for _ in range(2):
          var3.apply_gradients(opt3, torch.from_numpy(gk3).to(dev), torch.from_numpy(gg3).to(dev))
        got3 = var3.lookup(torch.from_numpy(gk3).to(dev)).cpu().numpy()
        f32 = np.float32
        gsum3 = f32(0)
        for r3 in range(world):                      # summed on the owner in source-rank order
          gsum3 = f32(gsum3 + f32(0.01 * (r3 + 1)))
        acc3, p3 = f32(0.1), f32(0)
        for _ in range(2):
          acc3 = f32(acc3 + f32(gsum3 * gsum3))
          p3 = f32(p3 - f32(f32(f32(0.1) * gsum3) / f32(np.sqrt(acc3))))
        if not np.array_equal(got3, np.full_like(got3, p3)):
          ok, msg = False, "owner-side sharded adagrad: %r vs %r" % (got3[0, 0], p3)
        torch.cuda.synchronize()
        dist.barrier()
        # coordinated growth: every rank re-creates its shard in a bigger symmetric region and streams its rows AND its
        # optimizer slots across; lookups, sizes and the next optimizer step must not notice
        size_before = int(var3.local.size())
        if var3.maybe_grow(threshold=0.99):
          ok, msg = False, "maybe_grow grew a table at load %.3f" % var3.load()
        var3.grow(1 << 19)
        if var3.capacity != 1 << 19 or int(var3.local.size()) != size_before:
          ok, msg = False, "grow: capacity %d size %d (was %d)" % (var3.capacity, int(var3.local.size()), size_before)
        if not np.array_equal(var3.lookup(torch.from_numpy(gk3).to(dev)).cpu().numpy(), got3):
          ok, msg = False, "grow: rows changed"
        var3.apply_gradients(opt3, torch.from_numpy(gk3).to(dev), torch.from_numpy(gg3).to(dev))   # accumulators survived
        acc3 = f32(acc3 + f32(gsum3 * gsum3))
        p3 = f32(p3 - f32(f32(f32(0.1) * gsum3) / f32(np.sqrt(acc3))))
        got4 = var3.lookup(torch.from_numpy(gk3).to(dev)).cpu().numpy()
        if not np.array_equal(got4, np.full_like(got4, p3)):
          ok, msg = False, "step after grow: %r vs %r" % (got4[0, 0], p3)
        torch.cuda.synchronize()
        dist.barrier()
      if var.tables[0].stats()["error_flags"] != 0:
        ok, msg = False, "error flags in mode %s" % mode
  except Exception as e:  # noqa: BLE001
    ok, msg = False, repr(e)
  q.put((rank, ok, msg))
  dist.barrier()
  dist.destroy_process_group()


def test_two_gpu_sharded_table():
  import torch
  import torch.multiprocessing as mp
  if torch.cuda.device_count() < 2:
    pytest.skip("needs >= 2 GPUs")
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29700 + (os.getpid() % 200)
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=300) for _ in range(world)]
  for p in procs:
    p.join(timeout=120)
  assert all(ok for _, ok, _ in res), res
