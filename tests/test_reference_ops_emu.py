"""The reference's op-level tests (tensorflow_recommenders_addons/dynamic_embedding/python/kernel_tests/
dynamic_embedding_ops_test.py), restated with torch tensors and run END TO END on the CPU: the Python mirror over the
EMULATED libdetable (tests/emu/backend.py -- the engine's own kernels and host code executed by the SIMT emulator).
The same API calls run on the GPU in tests/test_table_gpu.py / tests/test_fused_gpu.py; here every reference case is
covered, including the sharded ("devices * n", :329,392) and 3-D sparse ones.  Reference line numbers are in the
docstrings; tolerances are the reference's (assertAllClose = 1e-6)."""
import math

import numpy as np
import pytest
import torch

from tests.emu import backend


@pytest.fixture(autouse=True)
def _emu():
  with backend.installed():
    yield


def _de():
  from recommenders_addons_b200 import dynamic_embedding as de
  return de


def _truncated_normal(stddev, seed):
  g = torch.Generator().manual_seed(seed)

  def init(shape):
    x = torch.randn(list(shape), generator=g)
    while bool((x.abs() > 2).any()):   # tf.truncated_normal: redraw beyond two standard deviations
      x = torch.where(x.abs() > 2, torch.randn(list(shape), generator=g), x)
    return x * stddev
  return init


def _random_weights(name, vocab_size=4, embed_dim=4, num_shards=1):
  """:252-275"""
  return _de().get_variable(name, devices=["cpu"] * num_shards, dim=embed_dim,
                            initializer=_truncated_normal(1.0 / math.sqrt(vocab_size), 1))


def _ids_and_weights_2d(embed_dim=4):
  """:187-216"""
  de = _de()
  ind = torch.tensor([[0, 0], [0, 1], [0, 2], [1, 0], [3, 0], [4, 0], [4, 1]])
  ids = torch.tensor([0, 1, -100, -100, 2, 0, 1])
  w = torch.tensor([1.0, 2.0, 1.0, 1.0, 3.0, 0.0, -0.5])
  return de.SparseIds(ind, ids, (5, embed_dim)), de.SparseIds(ind, w, (5, embed_dim))


def _ids_and_weights_3d(embed_dim=4):
  """:219-250"""
  de = _de()
  ind = torch.tensor([[0, 0, 0], [0, 0, 1], [0, 0, 2], [0, 1, 0], [1, 0, 0], [1, 1, 0], [1, 1, 1]])
  ids = torch.tensor([0, 1, -100, -100, 2, 0, 1])
  w = torch.tensor([1.0, 2.0, 1.0, 1.0, 3.0, 0.0, -0.5])
  return de.SparseIds(ind, ids, (2, 3, embed_dim)), de.SparseIds(ind, w, (2, 3, embed_dim))


def _init(var, valid_ids):
  """the tests' "# init" block: look the ids up (initializer rows) and write them back"""
  ids = torch.tensor(valid_ids)
  vals = var.lookup(ids)
  var.upsert(ids, vals)
  return vals.numpy()


def _close(a, b):
  np.testing.assert_allclose(np.asarray(a, dtype=np.float32), np.asarray(b, dtype=np.float32), rtol=1e-6, atol=1e-6)


Z = [0.0] * 4


# ---- EmbeddingLookupTest (:322-800) ---------------------------------------------------------------------------------
def test_simple_sharded():
  """:324-349"""
  de = _de()
  emb = de.get_variable("t300", devices=["cpu"] * 2, initializer=2.0)
  ids = torch.tensor([0, 1, 2, 3, 4])
  embedding, trainable = de.embedding_lookup(emb, ids, max_norm=1.0, return_trainable=True)
  _close(embedding.detach(), [[1.0]] * 5)
  trainable.update_op()
  assert int(emb.size()) == 5 and int(emb.size(0)) == 3 and int(emb.size(1)) == 2


def test_max_norm():
  """:351-361"""
  de = _de()
  emb = de.get_variable("t310", initializer=2.0, devices=["cpu"])
  assert de.embedding_lookup(emb, torch.tensor([0]), max_norm=1.0).tolist() == [[1.0]]


def test_max_norm_nontrivial():
  """:363-380"""
  de = _de()
  emb = de.get_variable("t320", initializer=2.0, dim=2, devices=["cpu"])
  ids = torch.tensor([0, 1])
  emb.upsert(ids, torch.tensor([[2.0, 4.0], [3.0, 1.0]]))
  no_norm = de.embedding_lookup(emb, ids)
  got = de.embedding_lookup(emb, ids, max_norm=2.0)
  norms = (no_norm * no_norm).sum(1).sqrt()
  _close(got, 2 * no_norm / torch.stack([norms, norms], 1))


def test_sharded_custom_partitioner_int32_ids():
  """:382-408"""
  de = _de()
  emb = de.get_variable("t330", partitioner=lambda keys, shard_num: (keys % 2).to(torch.int32), devices=["cpu"] * 3,
                        initializer=2.0)
  emb.upsert(torch.tensor([0, 1, 2, 3, 4]), torch.tensor([[0.0], [1.0], [2.0], [3.0], [4.0]]))
  got = de.embedding_lookup(emb, torch.tensor([1, 3, 2, 3, 0]))
  _close(got, [[1.0], [3.0], [2.0], [3.0], [0.0]])
  assert tuple(got.shape) == (5, 1)
  assert [int(emb.size(i)) for i in range(3)] == [3, 2, 0]


def test_sharded_multi_lookup_on_one_variable():
  """:410-440"""
  de = _de()
  emb = de.get_variable("t340", devices=["cpu"] * 3, initializer=2.0)
  ids = torch.tensor([0, 1, 2, 3, 4])
  emb.upsert(ids, torch.tensor([[0.0], [1.0], [2.0], [3.0], [4.0]]))
  _close(de.embedding_lookup(emb, torch.tensor([1, 3, 2])), [[1.0], [3.0], [2.0]])
  _close(de.embedding_lookup(emb, torch.tensor([3, 4])), [[3.0], [4.0]])
  emb.upsert(ids, torch.tensor([[10.0], [11.0], [12.0], [13.0], [14.0]]))
  _close(de.embedding_lookup(emb, torch.tensor([3, 4])), [[13.0], [14.0]])


def test_higher_rank_and_shapes():
  """:442-464 and test_embedding_lookup_shape :746-798: ids of any shape -> ids.shape + [dim]"""
  de = _de()
  rng = np.random.default_rng(8)
  for dim in (1, 10):
    params = de.get_variable("t350-%d" % dim, initializer=2.0, dim=dim, devices=["cpu"])
    for ids_shape in ([3, 2], [4, 3], [4, 3, 10]):
      ids = torch.from_numpy(rng.integers(0, 2**31, size=ids_shape, dtype=np.int64))
      simple = params.lookup(ids)
      params.upsert(ids.reshape(-1), simple.reshape(-1, dim))
      got = de.embedding_lookup(params, ids)
      assert torch.equal(simple, got) and list(got.shape) == ids_shape + [dim]


def test_type_checks():
  """test_safe_embedding_lookup_sparse_inconsistent_ids_type / _weights_type :1167-1203, static checks :466-544"""
  de = _de()
  var = _random_weights("typecheck")
  sp, sw = _ids_and_weights_2d()
  with pytest.raises(TypeError):
    de.safe_embedding_lookup_sparse(var, de.SparseIds(sp.indices, sp.values.to(torch.int32), sp.dense_shape), sw)
  with pytest.raises(TypeError):
    de.safe_embedding_lookup_sparse(var, sp, de.SparseIds(sw.indices, sw.values.to(torch.float16), sw.dense_shape))
  with pytest.raises(TypeError):
    de.embedding_lookup_sparse(var, de.SparseIds(sp.indices, sp.values.to(torch.int32), sp.dense_shape), None)
  with pytest.raises(TypeError):
    de.embedding_lookup(var, torch.tensor([1], dtype=torch.int32))
  with pytest.raises(TypeError):
    de.embedding_lookup(object(), torch.tensor([1]))


def test_dynamic_embedding_variable_clear():
  """:1414-1436"""
  de = _de()
  table = de.get_variable("t160", value_dtype=torch.int32, initializer=-1, devices=["cpu"])
  table.upsert(torch.tensor([[0, 1], [2, 3]]), torch.tensor([[[0], [1]], [[2], [3]]], dtype=torch.int32))
  assert int(table.size()) == 4
  table.clear()
  assert int(table.size()) == 0
  assert table.lookup(torch.tensor([0, 1, 3, 4])).tolist() == [[-1], [-1], [-1], [-1]]


# ---- EmbeddingLookupUniqueTest (:801-822) --------------------------------------------------------------------------
def test_embedding_lookup_unique():
  de = _de()
  dim, n = 5, 10
  emb = de.get_variable("t_unique_001", dim=dim, devices=["cpu"])
  rng = np.random.default_rng(0)
  table = rng.standard_normal((n, dim)).astype(np.float32)
  ids = rng.integers(0, n, (2, 3, 4))
  emb.upsert(torch.arange(n), torch.from_numpy(table))
  got = de.embedding_lookup_unique(emb, torch.from_numpy(ids))
  assert tuple(got.shape) == (2, 3, 4, dim)
  np.testing.assert_array_equal(got.numpy(), table[ids])


# ---- SafeEmbeddingLookupSparseTest (:992-1412) ------------------------------------------------------------------------
def test_safe_embedding_lookup_sparse_return_zero_vector():
  """:1007-1048"""
  de = _de()
  var = _random_weights("safe-zero")
  sp, sw = _ids_and_weights_2d()
  v = _init(var, [0, 1, 2, -100])
  _close(de.safe_embedding_lookup_sparse(var, sp, sw), [(1.0 * v[0] + 2.0 * v[1] + 1.0 * v[3]) / 4.0, v[3] * 1.0, Z, v[2], Z])


def test_safe_embedding_lookup_sparse_return_special_vector():
  """:1052-1087 (default_id=3)"""
  de = _de()
  var = _random_weights("safe-special")
  sp, sw = _ids_and_weights_2d()
  v = _init(var, [0, 1, 2, 3, -100])
  _close(de.safe_embedding_lookup_sparse(var, sp, sw, default_id=3),
         [(1.0 * v[0] + 2.0 * v[1] + 1.0 * v[4]) / 4.0, v[4], v[3], v[2], v[3]])


@pytest.mark.parametrize("num_shards", [1, 3])
def test_safe_embedding_lookup_sparse_no_weights_and_partitioned(num_shards):
  """:1091-1125 and (3 shards) :1129-1163"""
  de = _de()
  var = _random_weights("safe-now-%d" % num_shards, num_shards=num_shards)
  sp, _ = _ids_and_weights_2d()
  v = _init(var, [0, 1, 2, -100])
  _close(de.safe_embedding_lookup_sparse(var, sp, None), [(v[0] + v[1] + v[3]) / 3.0, v[3], Z, v[2], (v[0] + v[1]) / 2.0])


def test_safe_embedding_lookup_sparse_3d_return_zero_vector():
  """:1205-1229"""
  de = _de()
  var = _random_weights("safe3-zero")
  sp, sw = _ids_and_weights_3d()
  v = _init(var, [0, 1, 2, -100])
  _close(de.safe_embedding_lookup_sparse(var, sp, sw),
         [[(1.0 * v[0] + 2.0 * v[1] + 1.0 * v[3]) / 4.0, v[3], Z], [v[2], Z, Z]])


def test_safe_embedding_lookup_sparse_3d_return_special_vector():
  """:1232-1257 (default_id=3)"""
  de = _de()
  var = _random_weights("safe3-special")
  sp, sw = _ids_and_weights_3d()
  v = _init(var, [0, 1, 2, 3, -100])
  _close(de.safe_embedding_lookup_sparse(var, sp, sw, default_id=3),
         [[(1.0 * v[0] + 2.0 * v[1] + 1.0 * v[4]) / 4.0, v[4], v[3]], [v[2], v[3], v[3]]])


@pytest.mark.parametrize("num_shards", [1, 3])
def test_safe_embedding_lookup_sparse_3d_no_weights_and_partitioned(num_shards):
  """:1260-1286 and (3 shards) :1289-1323"""
  de = _de()
  var = _random_weights("safe3-now-%d" % num_shards, num_shards=num_shards)
  sp, _ = _ids_and_weights_3d()
  v = _init(var, [0, 1, 2, -100])
  _close(de.safe_embedding_lookup_sparse(var, sp, None),
         [[(v[0] + v[1] + v[3]) / 3.0, v[3], Z], [v[2], (v[0] + v[1]) / 2.0, Z]])


def test_safe_embedding_lookup_sparse_with_initializer():
  """:1326-1390 (scaled down 16x): a random initializer gives every missing id its OWN row -- mean / stddev of the
  combined result match the reference's targets for its 50 % fill (0.0, 0.00029), rtol = atol = 2e-4"""
  de = _de()
  embed_dim, shape = 8, (16, 64, 32)
  total = shape[0] * shape[1] * shape[2]
  rng = np.random.default_rng(0)
  g = torch.Generator().manual_seed(5)
  var = de.get_variable("safe-init-bugfix", devices=["cpu"] * 3, dim=embed_dim,
                        initializer=lambda s: torch.randn(list(s), generator=g) * 0.001)
  flat = np.unique(rng.integers(0, total, int(total * 0.5)))
  ind = np.stack([flat // (shape[1] * shape[2]), (flat % (shape[1] * shape[2])) // shape[2], flat % shape[2]], 1)
  ids = rng.integers(-0x7FFFFFFFFFFFFFFF, 0x7FFFFFFFFFFFFFFF, flat.size, dtype=np.int64)
  sp = de.SparseIds(torch.from_numpy(ind), torch.from_numpy(ids), shape)
  vals = de.safe_embedding_lookup_sparse(var, sp, None, combiner="mean").numpy()
  assert vals.shape == (shape[0], shape[1], embed_dim)
  assert vals[0][0][0] != vals[0][0][1]
  np.testing.assert_allclose(vals.mean(), 0.0, rtol=2e-4, atol=2e-4)
  np.testing.assert_allclose(vals.std(), 0.00029, rtol=2e-4, atol=2e-4)


def test_embedding_lookup_sparse_with_initializer():
  """:702-744: embedding_lookup_sparse on a variable whose ids are all missing: per-id initializer rows, not one row"""
  de = _de()
  g = torch.Generator().manual_seed(6)
  var = de.get_variable("sp-init", dim=8, devices=["cpu"], initializer=lambda s: torch.randn(list(s), generator=g) * 0.001)
  n = 4096
  ind = torch.stack([torch.arange(n), torch.zeros(n, dtype=torch.int64)], 1)
  sp = de.SparseIds(ind, torch.arange(n) * 7919 + 1, (n, 1))
  vals = de.embedding_lookup_sparse(var, sp, None, combiner="sum").numpy()
  assert vals.shape == (n, 8) and len(np.unique(vals[:, 0])) > n // 2
  np.testing.assert_allclose(vals.std(), 0.001, rtol=0.05)


def test_embedding_lookup_sparse_max_norm_fused_and_composed_agree():
  """embedding_lookup_sparse(..., max_norm): the fused kernel (rows of up to 32 vectors) and the composed path (wider
  rows, trainable lookups) give the same result; max_norm clips every looked-up row before it is weighted"""
  de = _de()
  from recommenders_addons_b200.dynamic_embedding import ops
  rng = np.random.default_rng(4)
  for dim in (8, 200):
    var = de.get_variable("mn-%d" % dim, dim=dim, initializer=0.0, devices=["cpu"])
    keys = torch.arange(50)
    rows = torch.from_numpy(rng.normal(0, 1.0 / np.sqrt(dim), (50, dim)).astype(np.float32)) * torch.linspace(0.2, 3, 50)[:, None]
    var.upsert(keys, rows)
    n = 120
    ind = torch.stack([torch.sort(torch.from_numpy(rng.integers(0, 30, n))).values, torch.arange(n)], 1)
    sp = de.SparseIds(ind, torch.from_numpy(rng.integers(0, 60, n)), (30, n))
    sw = de.SparseIds(ind, torch.from_numpy(rng.uniform(0.5, 2, n).astype(np.float32)), (30, n))
    assert ops._clip_fusable(dim) == (dim == 8)
    got = de.embedding_lookup_sparse(var, sp, sw, combiner="mean", max_norm=1.0)
    composed, _ = de.embedding_lookup_sparse(var, sp, sw, combiner="mean", max_norm=1.0, return_trainable=True)
    _close(got, composed.detach())
    clipped = rows * (1.0 / torch.maximum(rows.norm(dim=1, keepdim=True), torch.tensor(1.0)))
    var2 = de.get_variable("mn-ref-%d" % dim, dim=dim, initializer=0.0, devices=["cpu"])
    var2.upsert(keys, clipped)
    _close(got, de.embedding_lookup_sparse(var2, sp, sw, combiner="mean"))


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("use_w", [False, True])
@pytest.mark.parametrize("dim", [4, 5, 64, 132])
def test_trainable_sparse_lookup_runs_the_fused_segment_sum_and_its_backward(combiner, use_w, dim):
  """training flavour of embedding_lookup_sparse (dynamic_embedding_ops.py:247-289; optimizer tests :747-1005): the
  forward over the TrainableWrapper's dense rows is det_sparse_segment_sum -- BIT-identical to the sequential oracle --
  and the gradient of the rows equals autograd through the plain torch restatement (float64) to 1e-6"""
  from oracle import oracle as O
  from recommenders_addons_b200.dynamic_embedding.ops import sparse_segment_sum_rows
  rng = np.random.default_rng(dim * 3 + len(combiner) + use_w)
  U, batch = 37, 23
  lens = rng.integers(0, 7, size=batch)
  seg = np.repeat(np.arange(batch), lens).astype(np.int32)
  idx = rng.integers(0, U, size=seg.shape[0]).astype(np.int32)
  w = rng.uniform(0.25, 2.0, size=seg.shape[0]).astype(np.float32) if use_w else None
  rows = rng.normal(0, 0.5, (U, dim)).astype(np.float32)
  t_rows = torch.tensor(rows, requires_grad=True)
  out = sparse_segment_sum_rows(t_rows, torch.tensor(idx), torch.tensor(seg), None if w is None else torch.tensor(w), batch, combiner)
  ot = O.PortTable(dim)
  ot.insert(np.arange(U, dtype=np.int64), rows)
  exp = O.embedding_lookup_sparse(ot, idx.astype(np.int64), seg, w, batch, combiner)
  np.testing.assert_array_equal(out.detach().numpy(), exp)
  gout = torch.tensor(rng.normal(0, 1, (batch, dim)).astype(np.float32))
  out.backward(gout)
  # reference gradient: autograd through gather * w -> index_add -> normalise, in float64
  r64 = torch.tensor(rows, dtype=torch.float64, requires_grad=True)
  w64 = torch.ones(seg.shape[0], dtype=torch.float64) if w is None else torch.tensor(w, dtype=torch.float64)
  s64 = torch.tensor(seg, dtype=torch.int64)
  o64 = torch.zeros((batch, dim), dtype=torch.float64).index_add(0, s64, r64[torch.tensor(idx, dtype=torch.int64)] * w64[:, None])
  if combiner != "sum":
    den = torch.zeros(batch, dtype=torch.float64).index_add(0, s64, w64 if combiner == "mean" else w64 * w64)
    den = den.sqrt() if combiner == "sqrtn" else den
    o64 = torch.where((den > 0)[:, None], o64 / den.clamp_min(1e-300)[:, None], o64)
  o64.backward(gout.double())
  np.testing.assert_allclose(t_rows.grad.numpy(), r64.grad.numpy(), rtol=1e-5, atol=1e-6)


def test_embedding_lookup_sparse_trainable_path_end_to_end(monkeypatch):
  """de.embedding_lookup_sparse(..., return_trainable=True) with DET_SPARSE_TRAIN_FUSED=1: values match the forward-only
  fused kernel bit for bit and a gradient reaches the TrainableWrapper's rows"""
  monkeypatch.setenv("DET_SPARSE_TRAIN_FUSED", "1")
  de = _de()
  var = de.get_variable("sparse-train-e2e", devices=["cpu"], dim=8, initializer=0.0)
  keys = torch.arange(0, 50)
  g = torch.Generator().manual_seed(3)
  var.upsert(keys, torch.randn(50, 8, generator=g))
  ids = torch.tensor([3, 7, 3, 11, 49, 7, 7, 60])
  indices = torch.tensor([[0, 0], [0, 1], [1, 0], [3, 0], [3, 1], [3, 2], [4, 0], [4, 1]])
  sp = de.SparseIds(indices, ids, (5, 3))
  spw = de.SparseIds(indices, torch.tensor([1.0, 2.0, 0.5, 1.0, 1.0, 3.0, 2.0, 1.0]), (5, 3))
  fwd = de.embedding_lookup_sparse(var, sp, spw, combiner="mean")
  out, tw = de.embedding_lookup_sparse(var, sp, spw, combiner="mean", return_trainable=True)
  assert torch.equal(out.detach(), fwd)
  out.sum().backward()
  assert tw.values.grad is not None and tw.values.grad.shape == tw.values.shape and bool(tw.values.grad.abs().sum() > 0)
