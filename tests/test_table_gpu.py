"""Parity of the CUDA table (through the Python mirror -> C ABI) with the oracle: the reference's
known-answer tests, the committed golden streams generated from the reference's own libcuckoo, and
randomized op streams checked against oracle/cuckoo_port.c.  Integer/key/exists results are bit-exact;
fp32 rows are copies or single IEEE adds, so they are compared bit-exactly too."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import GpuTableNp, golden_files, replay_golden, sorted_export

pytestmark = pytest.mark.gpu

GPU_DIMS = [1, 2, 4, 8, 10, 16, 32, 64, 100, 200]  # dynamic_embedding_variable_test.py:397


def _torch():
  import torch
  return torch


@pytest.mark.parametrize("dim", GPU_DIMS)
def test_variable_known_answer(dim):
  """dynamic_embedding_variable_test.py:394-468."""
  t = GpuTableNp(dim)
  assert t.size() == 0
  t.insert([0, 1, 2, 3], np.array([[0] * dim, [1] * dim, [2] * dim, [3] * dim], np.float32))
  assert t.size() == 4
  t.remove([1, 5])
  assert t.size() == 3
  out = t.find([0, 1, 5], np.full(dim, -1, np.float32))
  np.testing.assert_array_equal(out, np.array([[0] * dim, [-1] * dim, [-1] * dim], np.float32))
  k, v = t.export()
  np.testing.assert_array_equal(np.sort(k), [0, 2, 3])
  np.testing.assert_array_equal(np.sort(v, axis=0), np.array([[0] * dim, [2] * dim, [3] * dim], np.float32))


@pytest.mark.parametrize("dtype", ["int32", "int64", "int8", "float16", "bfloat16", "float64"])
@pytest.mark.parametrize("dim", [1, 8, 10, 64])
def test_variable_known_answer_other_dtypes(dtype, dim):
  """kv_list of the GPU build: int64 x {float32,int32,half,int8,int64,bfloat16} (:398-401)."""
  torch = _torch()
  td = getattr(torch, dtype)
  t = GpuTableNp(dim, dtype=td)
  vals = torch.tensor([[0] * dim, [1] * dim, [2] * dim, [3] * dim]).to(td)
  t.t.insert(torch.tensor([0, 1, 2, 3]), vals)
  assert t.size() == 4
  t.remove([1, 5])
  assert t.size() == 3
  out = t.t.lookup(torch.tensor([0, 1, 5]), dynamic_default_values=torch.full((dim,), -1).to(td))
  assert torch.equal(out.cpu(), torch.tensor([[0] * dim, [-1] * dim, [-1] * dim]).to(td))
  k, v = t.t.export()
  order = torch.argsort(k)
  assert k[order].tolist() == [0, 2, 3]
  assert torch.equal(v[order].cpu(), torch.tensor([[0] * dim, [2] * dim, [3] * dim]).to(td))
  # accum in the table's own arithmetic: found & exist -> add ; absent & !exist -> insert
  t.t.accum(torch.tensor([0, 2, 7, 8]), torch.tensor([[5] * dim] * 4).to(td), torch.tensor([True, True, False, True]))
  out = t.t.lookup(torch.tensor([0, 2, 7, 8]), dynamic_default_values=torch.full((dim,), -1).to(td))
  assert torch.equal(out.cpu(), torch.tensor([[5] * dim, [7] * dim, [5] * dim, [-1] * dim]).to(td))
  assert t.size() == 4


@pytest.mark.parametrize("dim", GPU_DIMS)
def test_find_with_exists_and_accum_known_answer(dim):
  """dynamic_embedding_variable_test.py:470-563 through de.Variable.accum: {0->10, 2->2, 3->13, 100->99}."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  var = de.Variable(dim=dim, initializer=-1.0, name="taccum1-%d" % dim)
  dev = var.tables[0].device
  f = torch.float32
  var.upsert(torch.tensor([0, 1, 2, 3], device=dev),
             torch.tensor([[0] * dim, [1] * dim, [2] * dim, [3] * dim], dtype=f, device=dev))
  accum_keys = torch.tensor([0, 1, 100, 3], device=dev)
  _, exists = var.lookup(accum_keys, return_exists=True)
  assert exists.tolist() == [True, True, False, True]
  var.upsert(torch.tensor([100], device=dev), torch.tensor([[99] * dim], dtype=f, device=dev))
  var.remove(torch.tensor([1], device=dev))
  assert int(var.size()) == 4
  old = torch.tensor([[0] * dim, [1] * dim, [2] * dim, [3] * dim], dtype=f, device=dev)
  new = torch.tensor([[10] * dim, [11] * dim, [100] * dim, [13] * dim], dtype=f, device=dev)
  var.accum(accum_keys, old, new, exists)
  k, v = var.export()
  assert sorted(k.tolist()) == [0, 2, 3, 100]
  np.testing.assert_array_equal(np.sort(v.cpu().numpy(), axis=0),
                                np.array([[2] * dim, [10] * dim, [13] * dim, [99] * dim], np.float32))


@pytest.mark.parametrize("path", golden_files())
def test_golden_streams(path):
  """Fixtures generated from the reference's own libcuckoo (oracle/gen_golden.py)."""
  g = np.load(path)
  replay_golden(GpuTableNp(int(g["dim"])), g)


@pytest.mark.parametrize("dim,init_size", [(16, 0), (64, 0), (5, 64), (128, 1 << 16)])
def test_random_streams_vs_oracle(dim, init_size):
  """Same op stream into the CUDA table and the C port; starts at 8192 slots (or 64) so the table grows
  through several rehashes; removals leave tombstones that later inserts recycle."""
  rng = np.random.default_rng(dim * 7 + 1)
  gt, ot = GpuTableNp(dim, init_size=init_size), O.PortTable(dim)
  universe = rng.permutation(np.arange(-30000, 30000, dtype=np.int64))
  for it in range(28):
    n = int(rng.integers(1, 6000))
    keys = rng.choice(universe, size=n, replace=False)
    op = it % 4
    if op == 0 or it < 3:
      v = rng.normal(0, 0.01, (n, dim)).astype(np.float32)
      gt.insert(keys, v), ot.insert(keys, v)
    elif op == 1:
      d = rng.normal(0, 1, (n, dim)).astype(np.float32)
      a, ea = gt.find(keys, d, True)
      b, eb = ot.find(keys, d, True)
      np.testing.assert_array_equal(ea, eb)
      np.testing.assert_array_equal(a, b)
      d1 = rng.normal(0, 1, dim).astype(np.float32)
      np.testing.assert_array_equal(gt.find(keys, d1), ot.find(keys, d1))
    elif op == 2:
      ex = rng.integers(0, 2, n).astype(bool)
      v = rng.normal(0, 0.01, (n, dim)).astype(np.float32)
      gt.accum(keys, v, ex), ot.accum(keys, v, ex)
    else:
      gt.remove(keys[:n // 2]), ot.remove(keys[:n // 2])
    assert gt.size() == ot.size(), it
  ka, va = sorted_export(gt)
  kb, vb = sorted_export(ot)
  np.testing.assert_array_equal(ka, kb)
  np.testing.assert_array_equal(va, vb)
  st = gt.t.stats()
  assert st["error_flags"] == 0
  assert st["size"] == gt.size() and st["used_slots"] >= st["size"]
  if init_size <= 64:
    assert st["rehash_count"] >= 3


def test_special_keys_and_empty_inputs():
  """Every int64 is a legal key, including the two values the key plane uses as EMPTY / TOMBSTONE."""
  dim = 8
  lo = np.iinfo(np.int64).min
  gt, ot = GpuTableNp(dim), O.PortTable(dim)
  keys = np.array([lo, lo + 1, lo + 2, -1, 0, np.iinfo(np.int64).max], np.int64)
  v = np.arange(6 * dim, dtype=np.float32).reshape(6, dim)
  for t in (gt, ot):
    t.insert(keys, v)
    t.insert(keys[:2], v[:2] + 100)     # overwrite the special keys
    t.remove(keys[1:2])                 # remove one of them
    t.accum(keys[:3], np.ones((3, dim), np.float32), [True, False, True])  # add / re-insert / add
  assert gt.size() == ot.size() == 6
  d = np.full(dim, -7, np.float32)
  q = np.concatenate([keys, [5, lo + 3]])
  a, ea = gt.find(q, d, True)
  b, eb = ot.find(q, d, True)
  np.testing.assert_array_equal(ea, eb)
  np.testing.assert_array_equal(a, b)
  for x, y in zip(sorted_export(gt), sorted_export(ot)):
    np.testing.assert_array_equal(x, y)
  # empty inputs are no-ops (rank/shape cases, dynamic_embedding_variable_test.py:1572-1687)
  e = np.zeros(0, np.int64)
  gt.insert(e, np.zeros((0, dim), np.float32))
  gt.remove(e)
  assert gt.find(e, d).shape == (0, dim)
  assert gt.size() == 6
  gt.clear()
  assert gt.size() == 0
  assert gt.export()[0].shape[0] == 0
  assert not gt.find(keys, d, True)[1].any()


def test_tombstone_recycling_and_clear():
  dim = 4
  gt, ot = GpuTableNp(dim, init_size=1 << 14), O.PortTable(dim)
  rng = np.random.default_rng(5)
  keys = rng.permutation(1 << 20)[:9000].astype(np.int64)
  v = rng.normal(size=(9000, dim)).astype(np.float32)
  for rnd in range(6):
    gt.insert(keys, v + rnd), ot.insert(keys, v + rnd)
    kill = keys[rnd::3]
    gt.remove(kill), ot.remove(kill)
    assert gt.size() == ot.size()
  for x, y in zip(sorted_export(gt), sorted_export(ot)):
    np.testing.assert_array_equal(x, y)
  st = gt.t.stats()
  # 9000 live keys never exceed 0.75 * 16384: re-writing resident keys must not grow the table (the growth
  # decision counts the batch's really-new keys), and tombstones are recycled or purged in place
  assert st["capacity"] == 1 << 14 and st["error_flags"] == 0


def test_insert_repeat_data_keeps_size():
  """hkv_hashtable_ops_test.py:572-625 test_insert_repeat_data: 50k keys twice -> size stays 50k."""
  dim = 8
  gt = GpuTableNp(dim)
  keys = np.arange(50000, dtype=np.int64) * 7 - 3
  v = np.random.default_rng(0).normal(size=(50000, dim)).astype(np.float32)
  gt.insert(keys, v)
  assert gt.size() == 50000
  gt.insert(keys, v * 2)
  assert gt.size() == 50000
  np.testing.assert_array_equal(gt.find(keys, np.zeros(dim, np.float32)), v * 2)


def test_duplicate_keys_in_one_call_are_memory_safe():
  """Not part of the contract (the reference's GPU table requires unique keys,
  dynamic_embedding_variable.py:1377-1378) but must not corrupt the table: one slot per key."""
  dim = 16
  gt = GpuTableNp(dim)
  keys = np.repeat(np.arange(2000, dtype=np.int64), 5)
  rows = np.repeat(np.arange(10000, dtype=np.float32)[:, None], dim, 1)
  gt.insert(keys, rows)
  assert gt.size() == 2000
  k, v = sorted_export(gt)
  np.testing.assert_array_equal(k, np.arange(2000))
  cand = rows.reshape(2000, 5, dim)
  assert all((v[i, 0] == cand[i, :, 0]).any() for i in range(0, 2000, 97))


def test_sharded_variable_lookup_and_sizes():
  """EmbeddingLookupTest (dynamic_embedding_ops_test.py:324-440): shards faked by repeating one device;
  `devices * 2` with keys 0..4 gives per-shard sizes 3 / 2 under the (key & 0x7fffffff) % S rule."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  dim = 4
  var = de.Variable(dim=dim, devices=["cuda:0"] * 2, initializer=0.0, name="t-shard")
  dev = var.tables[0].device
  keys = torch.arange(5, device=dev)
  vals = torch.arange(5 * dim, dtype=torch.float32, device=dev).reshape(5, dim)
  var.upsert(keys, vals)
  assert int(var.size(0)) == 3 and int(var.size(1)) == 2 and int(var.size()) == 5
  ids = torch.tensor([[4, 0], [9, 3]], device=dev)
  out = de.embedding_lookup(var, ids)
  assert out.shape == (2, 2, dim)
  exp = torch.stack([vals[4], vals[0], torch.zeros(dim, device=dev), vals[3]]).reshape(2, 2, dim)
  assert torch.equal(out, exp)
  # 3 shards, random keys vs oracle, incl. return_exists and remove
  var3 = de.Variable(dim=dim, devices=["cuda:0"] * 3, initializer=-1.0, name="t-shard3")
  rng = np.random.default_rng(1)
  k = rng.choice(np.arange(-5000, 5000), 3000, replace=False).astype(np.int64)
  v = rng.normal(size=(3000, dim)).astype(np.float32)
  ot = O.PortTable(dim)
  ot.insert(k, v)
  var3.upsert(torch.from_numpy(k).to(dev), torch.from_numpy(v).to(dev))
  sizes = [int(var3.size(i)) for i in range(3)]
  owner = O.default_partition_fn(k, 3, True)
  assert sizes == [int((owner == i).sum()) for i in range(3)]
  q = rng.integers(-6000, 6000, 4000).astype(np.int64)
  got, ex = var3.lookup(torch.from_numpy(q).to(dev), return_exists=True)
  exp, eex = ot.find(q, np.full(dim, -1, np.float32), True)
  np.testing.assert_array_equal(ex.cpu().numpy(), eex)
  np.testing.assert_array_equal(got.cpu().numpy(), exp)
  var3.remove(torch.from_numpy(k[:500]).to(dev))
  ot.remove(k[:500])
  assert int(var3.size()) == ot.size()


def test_embedding_lookup_unique_and_initializer():
  """embedding_lookup_unique (dynamic_embedding_ops.py:64-117) and a callable (random) initializer giving one
  default row per looked-up key (dynamic_embedding_variable.py:919-931); lookup must NOT insert."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  dim = 8
  gen = torch.Generator(device="cuda").manual_seed(0)
  var = de.Variable(dim=dim, name="t-init",
                    initializer=lambda shape: torch.randn(shape, generator=gen, device="cuda") * 0.01)
  dev = var.tables[0].device
  ids = torch.tensor([[7, 7, 3], [3, 9, 7]], device=dev)
  out = de.embedding_lookup_unique(var, ids)
  assert out.shape == (2, 3, dim)
  assert torch.equal(out[0, 0], out[0, 1]) and torch.equal(out[0, 0], out[1, 2]) and torch.equal(out[0, 2], out[1, 0])
  assert not torch.equal(out[0, 0], out[0, 2])
  assert int(var.size()) == 0
  big = var.lookup(torch.arange(1 << 17, device=dev))
  assert abs(float(big.mean())) < 2e-4 and abs(float(big.std()) - 0.01) < 2e-4  # :565-588


def test_export_import_and_file_round_trip(tmp_path):
  """cuckoo_hashtable_ops_test.py:155-267: save_to_file_system / load_from_file_system are bit-exact on
  sorted keys AND values; file format = raw <name>-keys / <name>-values."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  dim = 16
  n = 20000
  rng = np.random.default_rng(2)
  k = rng.choice(np.arange(10**9), n, replace=False).astype(np.int64)
  v = rng.normal(size=(n, dim)).astype(np.float32)
  a = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * dim, name="ta")
  a.insert(torch.from_numpy(k), torch.from_numpy(v))
  a.save_to_file_system(str(tmp_path), file_name="ckpt", dirpath_env="__unset__")
  raw_k = np.fromfile(os.path.join(str(tmp_path), "ckpt-keys"), dtype=np.int64)
  raw_v = np.fromfile(os.path.join(str(tmp_path), "ckpt-values"), dtype=np.float32).reshape(-1, dim)
  assert raw_k.shape[0] == n
  o = np.argsort(raw_k)
  np.testing.assert_array_equal(raw_k[o], np.sort(k))
  np.testing.assert_array_equal(raw_v[o], v[np.argsort(k)])
  b = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * dim, name="tb")
  b.insert(torch.tensor([1, 2, 3]), torch.ones(3, dim))  # load = clear + insert
  b.load_from_file_system(str(tmp_path), file_name="ckpt", dirpath_env="__unset__")
  assert int(b.size()) == n
  kb, vb = b.export()
  ob = torch.argsort(kb)
  np.testing.assert_array_equal(kb[ob].cpu().numpy(), np.sort(k))
  np.testing.assert_array_equal(vb[ob].cpu().numpy(), v[np.argsort(k)])
  # A file pair laid out the way the reference's SaveToFileSystem writes it (cuckoo_hashtable_op.cc:310-392: the
  # table's dump -- here the REFERENCE ENGINE's own iteration order, from the oracle -- as raw little-endian int64 keys
  # and raw rows) loads into our table: same content, whatever the order.
  ref = O.best_table(dim)
  ref.insert(k, v)
  rk, rv = ref.export()
  rk.astype("<i8").tofile(str(tmp_path / "refckpt-keys"))
  rv.astype("<f4").tofile(str(tmp_path / "refckpt-values"))
  d = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * dim, name="td")
  d.load_from_file_system(str(tmp_path), file_name="refckpt", dirpath_env="__unset__")
  assert int(d.size()) == n
  kd, vd = d.export()
  od = torch.argsort(kd)
  np.testing.assert_array_equal(kd[od].cpu().numpy(), np.sort(k))
  np.testing.assert_array_equal(vd[od].cpu().numpy(), v[np.argsort(k)])
  # import / export of 168 keys (cuckoo_hashtable_ops_test.py:76-99)
  c = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * dim, name="tc")
  c.import_(torch.from_numpy(k[:168]), torch.from_numpy(v[:168]))  # :76-99 import/export of 168 keys
  kc, vc = c.export()
  assert kc.numel() == 168
  oc = torch.argsort(kc)
  np.testing.assert_array_equal(vc[oc].cpu().numpy(), v[:168][np.argsort(k[:168])])


def test_host_buffer_entry_points():
  """det_find_host / det_insert_host (pinned and pageable) == device-pointer path."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  dim = 64
  n = 300000  # several pipeline chunks
  t = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * dim, init_size=1 << 20)
  g = torch.Generator().manual_seed(0)
  keys = torch.randperm(1 << 22, generator=g)[:n]
  vals = torch.randn(n, dim, generator=g)
  for pin in (True, False):
    t.clear()
    kh, vh = (keys.pin_memory(), vals.pin_memory()) if pin else (keys, vals)
    t.insert_host(kh, vh)
    assert int(t.size()) == n
    q = torch.cat([keys[: n // 2], torch.arange(1 << 22, (1 << 22) + n // 2)])
    d = torch.full((dim,), -3.0)
    out = torch.empty(q.numel(), dim)
    ex = torch.empty(q.numel(), dtype=torch.bool)
    if pin:
      q, out, ex = q.pin_memory(), out.pin_memory(), ex.pin_memory()
    t.lookup_host(q, d, out, ex)
    ref, rex = t.lookup(q.cuda(), dynamic_default_values=d.cuda(), return_exists=True)
    assert torch.equal(out, ref.cpu()) and torch.equal(ex, rex.cpu())
    assert torch.equal(out[: n // 2], vals[: n // 2]) and bool((out[n // 2:] == -3.0).all())


def test_error_behaviour():
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  from recommenders_addons_b200._lib import DetError
  t = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * 4)
  with pytest.raises(TypeError, match="Signature mismatch. Keys must be dtype"):
    t.lookup(torch.tensor([1, 2], dtype=torch.int32))
  with pytest.raises(TypeError, match="Signature mismatch"):
    t.insert(torch.tensor([1, 2]), torch.zeros(2, 4, dtype=torch.float64))
  with pytest.raises(ValueError):
    t.insert(torch.tensor([1, 2]), torch.zeros(2, 5))
  small = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * 4, init_size=64, max_capacity=64)
  with pytest.raises(DetError, match="max_capacity"):
    small.insert(torch.arange(1000), torch.zeros(1000, 4))


def test_config_c1_flow_1m_keys_dim16():
  """BASELINE.json configs[0]: 1M int64 keys, dim 16 fp32: insert all -> find (50% hits) -> accum(exists from
  find) -> remove 10% -> export; CUDA == oracle bit for bit."""
  dim, n, batch = 16, 1 << 20, 65536
  rng = np.random.default_rng(42)
  keys = rng.permutation(np.arange(1, 4 * n, dtype=np.int64))[:2 * n]
  present, absent = keys[:n], keys[n:]
  vals = rng.normal(0, 0.01, (n, dim)).astype(np.float32)
  gt, ot = GpuTableNp(dim), O.best_table(dim, 0, threads=1)
  for b in range(0, n, batch):
    gt.insert(present[b:b + batch], vals[b:b + batch])
  ot.insert(present, vals)
  assert gt.size() == ot.size() == n
  q = np.concatenate([present[:n // 2], absent[:n // 2]])
  rng.shuffle(q)
  d = rng.normal(0, 0.01, (q.shape[0], dim)).astype(np.float32)
  a, ea = gt.find(q, d, True)
  b_, eb = ot.find(q, d, True)
  np.testing.assert_array_equal(ea, eb)
  np.testing.assert_array_equal(a, b_)
  delta = rng.normal(0, 0.01, (q.shape[0], dim)).astype(np.float32)
  gt.accum(q, delta, ea), ot.accum(q, delta, eb)
  gt.remove(present[:n // 10]), ot.remove(present[:n // 10])
  assert gt.size() == ot.size()
  ka, va = sorted_export(gt)
  kb, vb = sorted_export(ot)
  np.testing.assert_array_equal(ka, kb)
  np.testing.assert_array_equal(va, vb)


def test_async_host_entry_points_overlap_semantics():
  """det_find_host_async / det_insert_host_async (pinned buffers only): enqueue-and-return; lookups of one batch
  may be issued together with the write-back of another; det_host_sync drains both pipelines."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  from recommenders_addons_b200._lib import DetError
  dim, n = 64, 200000
  t = de.CuckooHashTable(torch.int64, torch.float32, [0.0] * dim, init_size=1 << 21)
  g = torch.Generator().manual_seed(3)
  ka = torch.randperm(1 << 22, generator=g)[:n].pin_memory()
  kb = (torch.randperm(1 << 22, generator=g)[:n] + (1 << 23)).pin_memory()  # disjoint from ka
  va, vb = torch.randn(n, dim, generator=g).pin_memory(), torch.randn(n, dim, generator=g).pin_memory()
  d = torch.full((dim,), -2.0).pin_memory()
  out = torch.empty(n, dim).pin_memory()
  ex = torch.empty(n, dtype=torch.bool).pin_memory()
  t.insert_host_async(ka, va)
  t.host_sync()
  t.insert_host_async(kb, vb)                 # write-back of batch b ...
  t.lookup_host_async(ka, d, out, ex)         # ... overlapped with the lookup of batch a
  t.host_sync()
  assert int(t.size()) == 2 * n
  assert bool(ex.all()) and torch.equal(out, va)
  t.lookup_host_async(kb, d, out, ex)
  t.host_sync()
  assert bool(ex.all()) and torch.equal(out, vb)
  with pytest.raises(DetError, match="pinned"):
    t.insert_host_async(torch.arange(10), torch.zeros(10, dim))


def test_variable_file_system_naming_and_reshard_on_load(tmp_path):
  """Variable.save_to_file_system writes the reference's `<name>_mht_<i>of<N>_rank<r>_size<s>-keys/-values`
  files (dynamic_embedding_variable.py:1041-1044); a variable with a DIFFERENT shard count restores them with
  load_from_file_system_with_restore_function (reshard-on-load, :360-450)."""
  import os
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  dim = 8
  rng = np.random.default_rng(11)
  k = rng.choice(np.arange(-10**6, 10**6), 30000, replace=False).astype(np.int64)
  v = rng.normal(size=(30000, dim)).astype(np.float32)
  a = de.Variable(dim=dim, devices=["cuda:0"] * 3, name="emb/ckpt_var")
  dev = a.tables[0].device
  a.upsert(torch.from_numpy(k).to(dev), torch.from_numpy(v).to(dev))
  a.save_to_file_system(str(tmp_path), proc_size=1, proc_rank=0, dirpath_env="__unset__")
  names = sorted(os.listdir(str(tmp_path)))
  assert names == sorted("emb_ckpt_var_mht_%dof3_rank0_size1-%s" % (i, s) for i in (1, 2, 3) for s in ("keys", "values"))
  # same topology: plain load
  a2 = de.Variable(dim=dim, devices=["cuda:0"] * 3, name="emb/ckpt_var")
  a2.load_from_file_system(str(tmp_path), dirpath_env="__unset__")
  assert [int(a2.size(i)) for i in range(3)] == [int(a.size(i)) for i in range(3)]
  # different topology: 2 shards read the 3 saved shards
  b = de.Variable(dim=dim, devices=["cuda:0"] * 2, name="emb/ckpt_var")
  b.load_from_file_system_with_restore_function(str(tmp_path))
  assert int(b.size()) == 30000
  owner = O.default_partition_fn(k, 2, True)
  assert [int(b.size(i)) for i in range(2)] == [int((owner == i).sum()) for i in range(2)]
  got = b.lookup(torch.from_numpy(k).to(dev))
  np.testing.assert_array_equal(got.cpu().numpy(), v)
