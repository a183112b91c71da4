"""One-sided sharded table (det_peer_*): ONE kernel probes the owner shard and moves rows directly.
Single-GPU coverage fakes the shards on one device (as the reference's tests do,
dynamic_embedding_ops_test.py:329); the real NVLink path is covered by tests/test_multigpu_gpu.py."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("dim", [4, 64])
def test_peer_fake_shards_vs_oracle(world, dim):
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  dev = torch.device("cuda", 0)
  shards = [de.Variable(dim=dim, init_size=1 << 16, initializer=-1.0, name="peer-%d-%d-%d" % (world, dim, i))
            for i in range(world)]
  pv = de.PeerShardedVariable(fake_shards=shards)
  rng = np.random.default_rng(world * 10 + dim)
  lo = np.iinfo(np.int64).min
  keys = rng.choice(np.arange(-40000, 40000), 20000, replace=False).astype(np.int64)
  keys[:3] = [lo, lo + 1, np.iinfo(np.int64).max]
  vals = rng.normal(0, 0.01, (keys.shape[0], dim)).astype(np.float32)
  ot = O.PortTable(dim)
  tk = lambda a: torch.from_numpy(a).to(dev)
  pv.upsert(tk(keys), tk(vals))
  ot.insert(keys, vals)
  # every key landed in the shard the reference's partition function names, sizes are exact
  owner = O.default_partition_fn(keys, world, True)
  assert [int(s.size()) for s in shards] == [int((owner == i).sum()) for i in range(world)]
  assert pv.size() == ot.size()
  for i, s in enumerate(shards):
    k, _ = s.export()
    assert bool((torch.from_numpy(O.default_partition_fn(k.cpu().numpy(), world, True)) == i).all())
  # lookups of present + absent keys, broadcast and full-size defaults
  q = np.concatenate([keys[::2], rng.integers(50000, 90000, 5000)]).astype(np.int64)
  rng.shuffle(q)
  got, ex = pv.lookup(tk(q), return_exists=True)
  exp, eex = ot.find(q, np.full(dim, -1, np.float32), True)
  np.testing.assert_array_equal(ex.cpu().numpy(), eex)
  np.testing.assert_array_equal(got.cpu().numpy(), exp)
  d = rng.normal(size=(q.shape[0], dim)).astype(np.float32)
  np.testing.assert_array_equal(pv.lookup(tk(q), default=tk(d)).cpu().numpy(), ot.find(q, d))
  # overwrite + new keys through the one-sided path; then the per-shard tables agree with plain lookups
  k2 = np.concatenate([keys[:5000], rng.integers(100000, 200000, 3000)]).astype(np.int64)
  k2 = np.unique(k2)
  v2 = rng.normal(0, 0.01, (k2.shape[0], dim)).astype(np.float32)
  pv.upsert(tk(k2), tk(v2))
  pv.phase_barrier()
  ot.insert(k2, v2)
  assert pv.size() == ot.size()
  allk = np.concatenate([keys, k2])
  np.testing.assert_array_equal(pv.lookup(tk(allk)).cpu().numpy(), ot.find(allk, np.full(dim, -1, np.float32)))
  for s in shards:
    assert s.tables[0].stats()["error_flags"] == 0
  pv.close()


def test_published_table_cannot_grow():
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  from recommenders_addons_b200._lib import DetError
  shards = [de.Variable(dim=4, init_size=1024, name="peer-fixed-%d" % i) for i in range(2)]
  pv = de.PeerShardedVariable(fake_shards=shards)
  with pytest.raises(DetError, match="max_capacity"):
    shards[0].upsert(torch.arange(0, 4000, 2, device="cuda"), torch.zeros(2000, 4, device="cuda"))
  pv.close()


def test_table_in_caller_provided_region():
  """det_table_create_in_region: the planes live in memory the caller owns (the symmetric-memory region of the
  multi-GPU path); behaviour is identical to an owned table but the capacity is fixed."""
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  from recommenders_addons_b200._lib import DetError
  from tests.helpers import golden_files, replay_golden, GpuTableNp
  g = np.load([p for p in golden_files() if p.endswith("dim16.npz")][0])
  dim, cap = 16, 1 << 14
  nbytes = de.CuckooHashTable.region_bytes(torch.float32, dim, cap, 1, 0)
  assert nbytes >= cap * (8 + 2 * dim * 4)
  region = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
  t = GpuTableNp.__new__(GpuTableNp)
  t.torch, t.dim, t.dtype = torch, dim, torch.float32
  t.t = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * dim, init_size=cap, num_slot_planes=1, region=region)
  t.dev = t.t.device
  replay_golden(t, g)
  assert t.t.capacity() == cap
  with pytest.raises(DetError, match="max_capacity|cannot grow"):
    t.t.insert(torch.arange(10**6, 10**6 + 20000), torch.zeros(20000, dim))
  with pytest.raises(DetError, match="region"):
    de.CuckooHashTable(torch.int64, torch.float32, [0.0] * dim, init_size=cap, region=region[: nbytes // 2])
  t.t.close()


@pytest.mark.parametrize("world", [2, 8])
def test_peer_route_inbox(world):
  """det_peer_route = partition + pack + send in one kernel: every (key, row) pair lands in the inbox of the shard
  the reference's partition function names, in its source's segment, with exact counts (fake shards on one GPU)."""
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  dim = 16
  dev = torch.device("cuda", 0)
  shards = [de.Variable(dim=dim, init_size=4096, name="route-%d-%d" % (world, i)) for i in range(world)]
  pv = de.PeerShardedVariable(fake_shards=shards)
  pv.attach_inbox(50000)
  rng = np.random.default_rng(world)
  for n in (1, 255, 40000):
    keys = rng.integers(-2**62, 2**62, n).astype(np.int64)
    rows = rng.normal(size=(n, dim)).astype(np.float32)
    pv.route(torch.from_numpy(keys).to(dev), torch.from_numpy(rows).to(dev))
    pv.phase_barrier()
    owner = O.default_partition_fn(keys, world, True)
    for sh in range(world):
      k, r, counts = pv.inbox_take(sh)
      assert counts[0] == int((owner == sh).sum()) and sum(counts[1:]) == 0   # this process routes as rank 0
      got = dict(zip(k.cpu().numpy().tolist(), map(tuple, r.cpu().numpy())))
      exp = dict(zip(keys[owner == sh].tolist(), map(tuple, rows[owner == sh])))
      assert got == exp
  for s in shards:
    assert s.tables[0].stats()["error_flags"] == 0
  pv.close()


def test_peer_apply_gradients_single_process():
  """route -> combine duplicate keys -> fused Adagrad on the owner shard == oracle tables stepped with the summed
  gradient (twin-model idea of dynamic_embedding_optimizer_test.py:349-440)."""
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  dim, world = 8, 1
  dev = torch.device("cuda", 0)
  shards = [de.Variable(dim=dim, init_size=1 << 15, initializer=0.0, num_slot_planes=1, name="route-opt")]
  pv = de.PeerShardedVariable(fake_shards=shards)
  pv.attach_inbox(20000)
  opt = de.FusedAdagrad(0.1, 0.1)
  p, a = O.PortTable(dim), O.PortTable(dim)
  rng = np.random.default_rng(0)
  for step in range(4):
    keys = rng.integers(0, 3000, 8000).astype(np.int64)          # duplicates: several "ranks" touch the same row
    g = rng.normal(0, 1e-2, (8000, dim)).astype(np.float32)
    pv.apply_gradients(opt, torch.from_numpy(keys).to(dev), torch.from_numpy(g).to(dev))
    # oracle: gradients of a key are summed in arrival order (route keeps no order -> compare with tolerance)
    uk, inv = np.unique(keys, return_inverse=True)
    gs = np.zeros((uk.shape[0], dim), np.float64)
    np.add.at(gs, inv, g.astype(np.float64))
    O.sparse_adagrad_step(p, a, uk, gs.astype(np.float32), 0.1, np.zeros(dim, np.float32), np.full(dim, 0.1, np.float32))
  k, v = shards[0].export()
  o = torch.argsort(k)
  ek, ev = p.export()
  eo = np.argsort(ek)
  np.testing.assert_array_equal(k[o].cpu().numpy(), ek[eo])
  np.testing.assert_allclose(v[o].cpu().numpy(), ev[eo], rtol=1e-5, atol=1e-7)
  pv.close()
