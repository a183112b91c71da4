"""Host-side logic that needs no GPU: partition function parity, import surface, and the multi-rank key
exchange (world_size 2, gloo) with an in-memory stand-in for the local shard."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as O


def test_import_surface():
  from recommenders_addons_b200 import dynamic_embedding as de
  for name in ("get_variable", "Variable", "embedding_lookup", "embedding_lookup_unique", "embedding_lookup_sparse",
               "safe_embedding_lookup_sparse", "DynamicEmbeddingOptimizer", "CuckooHashTable", "HkvHashTable",
               "CuckooHashTableCreator", "HkvHashTableCreator", "default_partition_fn"):
    assert hasattr(de, name), name


def test_default_partition_fn_matches_oracle():
  from recommenders_addons_b200 import dynamic_embedding as de
  rng = np.random.default_rng(0)
  k = rng.integers(-2**63, 2**63 - 1, size=4096, dtype=np.int64)
  for s in (1, 2, 3, 8):
    for gpu_mode in (True, False):
      got = de.default_partition_fn(torch.from_numpy(k), s, gpu_mode).numpy()
      np.testing.assert_array_equal(got, O.default_partition_fn(k, s, gpu_mode))


def test_table_requires_cuda_and_fails_loudly():
  from recommenders_addons_b200 import dynamic_embedding as de
  if torch.cuda.is_available():
    pytest.skip("GPU present")
  with pytest.raises(Exception):
    de.CuckooHashTable(torch.int64, torch.float32, [0.0] * 4, device="cpu")
  with pytest.raises(Exception):
    de.get_variable("no_gpu_var", dim=4)


def test_dtype_checks():
  from recommenders_addons_b200 import dynamic_embedding as de
  with pytest.raises(TypeError):
    de.CuckooHashTable(torch.int32, torch.float32, [0.0], device="cuda:0")


class _DictShard(object):
  """CPU stand-in with Variable's lookup/upsert surface, for exchange-logic tests only."""

  def __init__(self, dim):
    self.dim = dim
    self.d = {}

  def lookup(self, keys):
    out = torch.zeros((keys.numel(), self.dim))
    for i, k in enumerate(keys.tolist()):
      if k in self.d:
        out[i] = self.d[k]
    return out

  def upsert(self, keys, values):
    for k, v in zip(keys.tolist(), values):
      self.d[k] = v.clone()


def _cpu_partition(world):
  def f(keys):
    owner = torch.from_numpy(O.default_partition_fn(keys.numpy(), world, True)).long()
    perm = torch.sort(owner, stable=True).indices
    return keys[perm], perm.to(torch.int32), torch.bincount(owner, minlength=world)
  return f


def _worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from recommenders_addons_b200.dynamic_embedding.sharded import ShardedVariable
  dim = 4
  sv = ShardedVariable(_DictShard(dim), partition_impl=_cpu_partition(world),
                       gather_impl=lambda rows, perm: rows[perm.long()],
                       scatter_impl=lambda rows, perm: torch.empty_like(rows).index_copy_(0, perm.long(), rows))
  g = torch.Generator().manual_seed(100 + rank)
  keys = torch.unique(torch.randint(0, 10000, (500,), generator=g))
  vals = keys.float().reshape(-1, 1).repeat(1, dim) + 0.5 * rank  # value encodes (key, writer)
  sv.upsert(keys, vals)
  dist.barrier()
  # every key this rank owns must satisfy the partition rule
  own = torch.tensor(sorted(sv.local.d.keys()))
  assert bool((torch.from_numpy(O.default_partition_fn(own.numpy(), world, True)) == rank).all())
  # lookups of keys written by BOTH ranks come back in request order, from whichever rank owns them
  other = torch.unique(torch.randint(0, 10000, (500,), generator=torch.Generator().manual_seed(100 + (1 - rank))))
  q_keys = torch.cat([keys[:50], other[:50], torch.tensor([20001, 20002])])
  rows = sv.lookup(q_keys)
  base = rows[:, 0] - q_keys.float()
  ok = bool(((base == 0.0) | (base == 0.5) | (rows[:, 0] == 0)).all()) and bool((rows[-2:] == 0).all())
  ok = ok and bool((rows[:50, 0] - keys[:50].float()).abs().max() <= 0.5)
  q.put((rank, ok, len(sv.local.d)))
  dist.barrier()
  dist.destroy_process_group()


def test_sharded_exchange_world2_gloo():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + (os.getpid() % 500)
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in range(2)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert all(ok for _, ok, _ in res), res
  assert sum(n for _, _, n in res) > 0


def _fixture(indices, shape):
  from recommenders_addons_b200 import dynamic_embedding as de
  ids = torch.tensor([0, 1, -100, -100, 2, 0, 1])
  weights = torch.tensor([1.0, 2.0, 1.0, 1.0, 3.0, 0.0, -0.5])
  ind = torch.tensor(indices)
  return de.SparseIds(ind, ids, shape), de.SparseIds(ind, weights, shape)


def test_safe_lookup_preprocessing_2d_fixture():
  """_ids_and_weights_2d (dynamic_embedding_ops_test.py:187-216): row 0 valid ids + an invalid one, row 1 only an
  invalid id, row 2 empty, row 3 single id, row 4 only weights <= 0."""
  from recommenders_addons_b200.dynamic_embedding.ops import _safe_preprocess
  sp, sw = _fixture([[0, 0], [0, 1], [0, 2], [1, 0], [3, 0], [4, 0], [4, 1]], (5, 4))
  sp2, sw2, empty, shape = _safe_preprocess(sp, sw, "mean", None)
  # weights <= 0 pruned (row 4 becomes empty); rows 2 and 4 get one filler entry (id 0, weight 1)
  assert sp2.dense_shape == (5, 4) and shape == (5, 4)
  assert sp2.indices[:, 0].tolist() == [0, 0, 0, 1, 2, 3, 4]
  assert sp2.values.tolist() == [0, 1, -100, -100, 0, 2, 0]
  assert sw2.values.tolist() == [1.0, 2.0, 1.0, 1.0, 1.0, 3.0, 1.0]
  assert empty.tolist() == [2, 4]
  # combiner "sum" keeps non-positive weights; default_id fills the empty row
  sp3, sw3, empty3, _ = _safe_preprocess(sp, sw, "sum", 3)
  assert sp3.values.tolist() == [0, 1, -100, -100, 3, 2, 0, 1] and empty3.tolist() == [2]
  assert sw3.values.tolist() == [1.0, 2.0, 1.0, 1.0, 1.0, 3.0, 0.0, -0.5]
  # no weights at all: nothing is pruned
  sp4, sw4, empty4, _ = _safe_preprocess(sp, None, "mean", None)
  assert sw4 is None and sp4.values.numel() == 8 and empty4.tolist() == [2]


def test_safe_lookup_preprocessing_3d_fixture():
  """_ids_and_weights_3d (dynamic_embedding_ops_test.py:219-250): leading dims [2, 3] flatten to 6 rows."""
  from recommenders_addons_b200.dynamic_embedding.ops import _safe_preprocess
  sp, sw = _fixture([[0, 0, 0], [0, 0, 1], [0, 0, 2], [0, 1, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1]], (2, 3, 4))
  sp2, sw2, empty, shape = _safe_preprocess(sp, sw, "mean", None)
  assert sp2.dense_shape == (6, 4) and shape == (2, 3, 4)
  assert sp2.indices[:, 0].tolist() == [0, 0, 0, 1, 2, 3, 4, 5]
  assert empty.tolist() == [2, 4, 5]
  assert sp2.values.tolist() == [0, 1, -100, -100, 0, 2, 0, 0]


def test_sparse_combiners_cpu_oracle_against_reference_fixture():
  """SafeEmbeddingLookupSparseTest::test_safe_embedding_lookup_sparse_return_zero_vector semantics
  (dynamic_embedding_ops_test.py:1007-1040) on the oracle: mean of rows 0,1 weighted 1,2 (invalid id -> default 0)."""
  from recommenders_addons_b200.dynamic_embedding.ops import _safe_preprocess
  dim = 4
  emb = np.arange(3 * dim, dtype=np.float32).reshape(3, dim) + 1
  t = O.PortTable(dim)
  t.insert(np.array([0, 1, 2]), emb)
  sp, sw = _fixture([[0, 0], [0, 1], [0, 2], [1, 0], [3, 0], [4, 0], [4, 1]], (5, 4))
  sp2, sw2, empty, _ = _safe_preprocess(sp, sw, "mean", None)
  out = O.embedding_lookup_sparse(t, sp2.values.numpy(), sp2.indices[:, 0].numpy(), sw2.values.numpy(), 5, "mean")
  out[empty.numpy()] = 0
  exp = np.stack([(emb[0] * 1 + emb[1] * 2) / 4.0, np.zeros(dim), np.zeros(dim), emb[2], np.zeros(dim)]).astype(np.float32)
  np.testing.assert_allclose(out, exp, rtol=1e-6, atol=1e-6)


class _SgdStub(object):
  """optimizer stand-in: plain SGD on the dict shard (exchange-logic test only)."""
  iterations = 0

  def apply_sparse(self, shard, keys, grads):
    for k, g in zip(keys.tolist(), grads):
      shard.d[k] = shard.d.get(k, torch.zeros(shard.dim)) - 0.5 * g


def _cpu_unique(keys):
  u, idx = O.unique_first_occurrence(keys.numpy())
  return torch.from_numpy(u), torch.from_numpy(idx)


def _worker_bwd(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from recommenders_addons_b200.dynamic_embedding.sharded import ShardedVariable
  dim = 2
  sv = ShardedVariable(_DictShard(dim), partition_impl=_cpu_partition(world),
                       gather_impl=lambda rows, perm: rows[perm.long()],
                       scatter_impl=lambda rows, perm: torch.empty_like(rows).index_copy_(0, perm.long(), rows),
                       unique_impl=_cpu_unique)
  keys = torch.arange(0, 40)                      # BOTH ranks send gradients for the same 40 keys
  grads = torch.full((40, dim), float(rank + 1))
  sv.apply_gradients(_SgdStub(), keys, grads)
  dist.barrier()
  # the owner combined the gradients of both ranks before stepping: -0.5 * (1 + 2)
  ok = all(abs(float(v[0]) + 1.5) < 1e-6 for v in sv.local.d.values())
  owned = sorted(sv.local.d.keys())
  ok = ok and owned == [k for k in range(40) if (k & 0x7fffffff) % world == rank]
  q.put((rank, ok, len(owned)))
  dist.barrier()
  dist.destroy_process_group()


def test_sharded_backward_routes_and_combines_world2_gloo():
  """half-sync sparse update over the exchange (dynamic_embedding_optimizer.py:580-595): row gradients travel to
  the owner, duplicates from several ranks are summed there, sparse rows are never all-reduced."""
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29800 + (os.getpid() % 150)
  procs = [ctx.Process(target=_worker_bwd, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in range(2)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert all(ok for _, ok, _ in res), res
  assert sum(n for _, _, n in res) == 40


def _worker_agree(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from recommenders_addons_b200.dynamic_embedding.sharded import agree_on
  ok = agree_on(4096, what="max_items") == 4096
  try:
    agree_on(4096 + rank, what="max_items")           # a rank-local choice: refused on EVERY rank
    ok = False
  except ValueError as e:
    ok = ok and "min 4096, max 4097" in str(e)
  q.put((rank, ok))
  dist.barrier()
  dist.destroy_process_group()


def test_mailbox_capacity_must_be_agreed_world2_gloo():
  """attach_exchange / attach_inbox fix the segment offsets peers store to: every rank must pass the same capacity"""
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 30400 + (os.getpid() % 150)
  procs = [ctx.Process(target=_worker_agree, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in range(2)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert all(ok for _, ok in res), res


def _score_of(keys):
  return keys * 3 + 11          # CUSTOMIZED: one distinct score per key, so which keys an event evicts is fully determined


def _worker_evict(rank, world, port, q):
  """sharded table WITH eviction = the collective exchange over a local shard that has an eviction strategy (the reference:
  every Horovod rank owns one HkvHashTable, shadow_embedding_ops.py:397-447 over hkv_hashtable_ops.py).  The local shard
  is the real engine on the emulator; its twin is a second, unsharded table fed with exactly what the owner receives."""
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from tests.emu import backend
  with backend.installed():
    from recommenders_addons_b200 import dynamic_embedding as de
    from recommenders_addons_b200.dynamic_embedding.sharded import ShardedVariable
    dim, cap = 4, 512

    def shard(name):
      cfg = de.HkvHashTableConfig(init_capacity=cap, max_capacity=cap, evict_strategy=de.HkvEvictStrategy.CUSTOMIZED,
                                  gen_scores_fn=_score_of)
      return de.get_variable(name, key_dtype=torch.int64, value_dtype=torch.float32, devices=["cpu"], initializer=0.0,
                             dim=dim, init_size=cap, kv_creator=de.HkvHashTableCreator(config=cfg))

    local, twin = shard("evict_shard_%d" % rank), shard("evict_twin_%d" % rank)
    sv = ShardedVariable(local, partition_impl=_cpu_partition(world),
                         gather_impl=lambda rows, perm: rows[perm.long()],
                         scatter_impl=lambda rows, perm: torch.empty_like(rows).index_copy_(0, perm.long(), rows),
                         unique_impl=_cpu_unique)
    streams = []
    for step in range(6):                                   # 6 x 2 x 300 distinct keys >> 2 x 512 slots
      per_rank = [torch.arange(300, dtype=torch.int64) * 2 + r + 600 * step for r in range(world)]
      streams.append(per_rank)
    for per_rank in streams:
      keys = per_rank[rank]
      sv.upsert(keys, keys.float().reshape(-1, 1).repeat(1, dim))
      # what this owner received, in source-rank order (exchange_keys concatenates by source)
      got = torch.cat([k[torch.from_numpy(O.default_partition_fn(k.numpy(), world, True)).long() == rank] for k in per_rank])
      twin.upsert(got, got.float().reshape(-1, 1).repeat(1, dim))
    dist.barrier()
    n_local = int(local.size())
    lk, lv = local.export()
    tk, _ = twin.export()
    ok = n_local <= cap and sorted(lk.tolist()) == sorted(tk.tolist())
    ok = ok and bool((lv[:, 0] == lk.float()).all())                      # rows of the survivors are intact
    ok = ok and bool((torch.from_numpy(O.default_partition_fn(lk.numpy(), world, True)) == rank).all())
    # the survivors are the highest-scored keys this owner ever received
    seen = torch.cat([k for per_rank in streams for k in per_rank])
    mine = seen[torch.from_numpy(O.default_partition_fn(seen.numpy(), world, True)).long() == rank]
    ok = ok and sorted(lk.tolist()) == sorted(mine.tolist())[-n_local:]
    ok = ok and local.tables[0].stats()["evict_events"] > 0
    # a lookup through the exchange: survivors come back with their rows, evicted keys with the default (0)
    probe = torch.cat([seen[-50:], seen[:50]])
    rows = sv.lookup(probe)
    total = int(sv.size())
    alive = torch.isin(probe, torch.cat([lk, torch.tensor(q_peer(rank, world, lk), dtype=torch.int64)]))
    ok = ok and bool((rows[alive][:, 0] == probe[alive].float()).all()) and bool((rows[~alive] == 0).all())
    q.put((rank, ok, n_local, total))
  dist.barrier()
  dist.destroy_process_group()


def q_peer(rank, world, my_keys):
  """keys alive on the OTHER rank, exchanged with an all-gather of padded key lists"""
  n = torch.tensor([len(my_keys)])
  sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
  dist.all_gather(sizes, n)
  m = int(max(int(s) for s in sizes))
  pad = torch.full((m,), -1, dtype=torch.int64)
  pad[:len(my_keys)] = my_keys
  bufs = [torch.empty(m, dtype=torch.int64) for _ in range(world)]
  dist.all_gather(bufs, pad)
  return [k for r in range(world) if r != rank for k in bufs[r][:int(sizes[r])].tolist()]


def test_sharded_table_with_eviction_world2_gloo():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 30100 + (os.getpid() % 150)
  procs = [ctx.Process(target=_worker_evict, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=600) for _ in range(2)]
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  assert all(r[1] for r in res), res
  assert all(r[2] <= 512 for r in res) and res[0][3] == res[1][3] == sum(r[2] for r in res)


def _worker_a2a_layer(rank, world, port, q):
  """de.layers.AllToAllEmbedding with the reference's HkvHashTableCreator (an eviction strategy): the collective exchange
  over one evicting HkvHashTable per rank (HvdAllToAllEmbedding over hkv tables, embedding.py:545-595)."""
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  import numpy as np
  from tests.emu import backend
  with backend.installed():
    from recommenders_addons_b200 import dynamic_embedding as de
    dim, cap, lr = 4, 512, 0.5
    impls = dict(partition_impl=_cpu_partition(world), gather_impl=lambda rows, perm: rows[perm.long()],
                 scatter_impl=lambda rows, perm: torch.empty_like(rows).index_copy_(0, perm.long(), rows),
                 unique_impl=_cpu_unique)
    cfg = de.HkvHashTableConfig(init_capacity=cap, max_capacity=cap, evict_strategy=de.HkvEvictStrategy.LFU)
    layer = de.layers.AllToAllEmbedding(dim, cap, initializer=0.5, name="a2a_evict_%d" % rank, num_slot_planes=1,
                                        kv_creator=de.HkvHashTableCreator(config=cfg), devices=["cpu"], exchange_impls=impls)
    assert layer.collective and layer.evicting
    opt = de.FusedAdagrad(lr, 0.1)
    # phase 1 (no eviction yet): both ranks train overlapping ids; a sequential twin keeps params and accumulators
    par, acc = {}, {}
    f32 = np.float32
    ok = True
    for step in range(3):
      ids_of = [torch.tensor([[1, 2, 3, 2], [7, 1, 9 + r, 40 + step]], dtype=torch.int64) for r in range(world)]
      wts_of = [torch.tensor([[1., 2., -1., 4.], [2., 1., 1., -2.]]) * (r + 1) for r in range(world)]
      out = layer(ids_of[rank])
      exp = torch.tensor([[par.get(int(k), np.full(dim, 0.5, f32)) for k in row] for row in ids_of[rank].tolist()])
      ok = ok and bool(torch.equal(out.detach(), exp))
      (out * wts_of[rank].unsqueeze(-1)).sum().backward()
      layer.apply_gradients(opt)
      gsum = {}
      for r in range(world):                       # the owner adds the ranks' per-unique gradients in source order
        per = {}
        for k, w in zip(ids_of[r].reshape(-1).tolist(), wts_of[r].reshape(-1).tolist()):
          per[k] = f32(per.get(k, f32(0)) + f32(w))
        for k, g in per.items():
          gsum[k] = f32(gsum.get(k, f32(0)) + g)
      for k, g in gsum.items():
        gv = np.full(dim, g, f32)
        a1 = (acc.get(k, np.full(dim, 0.1, f32)) + gv * gv).astype(f32)
        par[k] = (par.get(k, np.full(dim, 0.5, f32)) - (f32(lr) * gv) / np.sqrt(a1)).astype(f32)
        acc[k] = a1
    probe = torch.tensor(sorted(par), dtype=torch.int64)
    layer.eval()
    got = layer(probe)
    ok = ok and bool(torch.equal(got, torch.tensor(np.stack([par[int(k)] for k in probe.tolist()]))))
    # phase 2: far more ids than 2 x 512 slots; the ids of phase 1 keep being trained (LFU keeps them)
    layer.train()
    hot = probe
    for step in range(8):
      cold = torch.arange(200, dtype=torch.int64) * world + rank + 1000 + 400 * step
      out = layer(torch.cat([hot, cold]))
      out.sum().backward()
      layer.apply_gradients(opt)
    local = layer.params.local
    n_local = int(local.size())
    ok = ok and n_local <= cap and local.tables[0].stats()["evict_events"] > 0
    lk, _ = local.export()
    mine_hot = hot[torch.from_numpy(O.default_partition_fn(hot.numpy(), world, True)).long() == rank]
    ok = ok and bool(torch.isin(mine_hot, lk).all())
    q.put((rank, ok, n_local, int(layer.params.size())))
  dist.barrier()
  dist.destroy_process_group()


def test_alltoall_embedding_with_an_evicting_kv_creator_world2_gloo():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 30700 + (os.getpid() % 150)
  procs = [ctx.Process(target=_worker_a2a_layer, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=600) for _ in range(2)]
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  assert all(r[1] for r in res), res
  assert res[0][3] == res[1][3] == sum(r[2] for r in res)


def test_math_and_data_flow_mirrors():
  """de.math / de.data_flow (python/ops/math_ops.py:60-215, data_flow_ops.py:40-61): the doc examples of the TF ops
  they stand for"""
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  c = torch.tensor([[1, 2, 3, 4], [-1, -2, -3, -4], [5, 6, 7, 8]])
  # tf.sparse.segment_sum examples
  assert de.math.sparse_segment_sum(c, [0, 1], [0, 0]).tolist() == [[0, 0, 0, 0]]
  assert de.math.sparse_segment_sum(c, [0, 1], [0, 1]).tolist() == [[1, 2, 3, 4], [-1, -2, -3, -4]]
  assert de.math.sparse_segment_sum(c, [0, 1, 2], [0, 0, 1]).tolist() == [[0, 0, 0, 0], [5, 6, 7, 8]]
  assert de.math.sparse_segment_sum(c, [0, 1], [0, 2], num_segments=4).tolist() == [[1, 2, 3, 4], [0] * 4, [-1, -2, -3, -4], [0] * 4]
  with pytest.raises(ValueError):
    de.math.sparse_segment_sum(c, [0, 1], [1, 0])
  # tf.sparse.fill_empty_rows example: rows 2 and 4 of a [5, 6] input are empty
  sp = de.SparseIds(torch.tensor([[0, 1], [0, 3], [1, 2], [1, 3], [3, 1], [3, 3]]), torch.tensor([1, 2, 3, 4, 5, 6]), (5, 6))
  filled, was_empty = de.math.sparse_fill_empty_rows(sp, 9)
  assert filled.indices.tolist() == [[0, 1], [0, 3], [1, 2], [1, 3], [2, 0], [3, 1], [3, 3], [4, 0]]
  assert filled.values.tolist() == [1, 2, 3, 4, 9, 5, 6, 9] and was_empty.tolist() == [False, False, True, False, True]
  # tf.sparse.reshape example: [2, 3, 6] -> [9, -1]
  sp = de.SparseIds(torch.tensor([[0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0], [1, 2, 3]]), torch.arange(5), (2, 3, 6))
  r = de.math.sparse_reshape(sp, [9, -1])
  assert r.dense_shape == (9, 4) and r.indices.tolist() == [[0, 0], [0, 1], [1, 2], [4, 2], [8, 1]]
  with pytest.raises(ValueError):
    de.math.sparse_reshape(sp, [7, -1])
  # tf.dynamic_partition / tf.dynamic_stitch examples
  parts = de.data_flow.dynamic_partition(torch.tensor([10, 20, 30, 40, 50]), torch.tensor([0, 0, 1, 1, 0]), 2)
  assert [p.tolist() for p in parts] == [[10, 20, 50], [30, 40]]
  idx = [torch.tensor(6), torch.tensor([4, 1]), torch.tensor([[5, 2], [0, 3]])]
  data = [torch.tensor([61, 62]), torch.tensor([[41, 42], [11, 12]]), torch.tensor([[[51, 52], [21, 22]], [[1, 2], [31, 32]]])]
  assert de.data_flow.dynamic_stitch(idx, data).tolist() == [[1, 2], [11, 12], [21, 22], [31, 32], [41, 42], [51, 52], [61, 62]]
  # partition + stitch round trip = the identity (make_partition / _stitch of de.Variable)
  x = torch.arange(12.).reshape(6, 2)
  p = torch.tensor([2, 0, 1, 0, 2, 1])
  pos = de.data_flow.dynamic_partition(torch.arange(6), p, 3)
  rows = de.data_flow.dynamic_partition(x, p, 3)
  assert torch.equal(de.data_flow.dynamic_stitch(pos, rows), x)
  assert de.data_flow.dynamic_stitch([torch.tensor([0, 0])], [torch.tensor([1., 2.])]).tolist() == [2.0]


def test_export_list_of_the_reference_is_importable():
  """tfra.dynamic_embedding.__all__ (dynamic_embedding/__init__.py:17-53): every hot-path name resolves; the names left
  out are the ones SURVEY.md 2 marks off the path (Redis tables, TF resource-variable wrappers, tf.train savers)"""
  from recommenders_addons_b200 import dynamic_embedding as de
  ref_all = ["CuckooHashTable", "CuckooHashTableConfig", "CuckooHashTableCreator", "HkvEvictStrategy", "HkvHashTable",
             "HkvHashTableConfig", "HkvHashTableCreator", "Variable", "TrainableWrapper", "DynamicEmbeddingOptimizer",
             "GraphKeys", "ModelMode", "RestrictPolicy", "TimestampRestrictPolicy", "FrequencyRestrictPolicy", "get_variable",
             "embedding_lookup", "embedding_lookup_sparse", "embedding_lookup_unique", "safe_embedding_lookup_sparse",
             "enable_inference_mode", "enable_train_mode", "get_model_mode", "trainable_wrapper_filter", "keras", "math",
             "data_flow", "shadow_ops"]
  assert [n for n in ref_all if not hasattr(de, n)] == []
  for n in ("Embedding", "BasicEmbedding", "FieldWiseEmbedding", "SquashedEmbedding", "HvdAllToAllEmbedding"):
    assert hasattr(de.keras.layers, n)
  out_of_scope = ["RedisTable", "RedisTableConfig", "RedisTableCreator", "DistributedVariableWrapper", "DEResourceVariable",
                  "train"]
  assert [n for n in out_of_scope if hasattr(de, n)] == []


def test_make_partition_like_the_reference():
  """dynamic_embedding_variable.py:131-154: data and positions split by partition index, in order; one shard = as is"""
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  data = torch.tensor([50, 60, 70, 80, 90])
  parts, idx = de.make_partition(data, torch.tensor([1, 0, 1, 2, 0]), 3)
  assert [p.tolist() for p in parts] == [[60, 90], [50, 70], [80]] and [i.tolist() for i in idx] == [[1, 4], [0, 2], [3]]
  parts, idx = de.make_partition(data, None, 1)
  assert parts[0] is data and idx is None
