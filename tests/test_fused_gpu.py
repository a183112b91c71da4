"""Parity of the fused kernels (K6 lookup_sparse, K7 Adagrad/Adam, K8 partition, unique) with the NumPy
restatements in oracle/oracle.py.  The kernels are compiled with --fmad=false and use IEEE sqrt/div, so
fp32 results are compared BIT-EXACTLY with the oracle; the reference's own tolerance (rtol=atol=1e-6,
dynamic_embedding_ops_test.py:875-969) is additionally checked against its test-oracle formulation."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import sorted_export

pytestmark = pytest.mark.gpu


def _torch():
  import torch
  return torch


@pytest.mark.parametrize("n", [1, 31, 1000, 70000])
def test_unique_first_occurrence(n):
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  rng = np.random.default_rng(n)
  ids = rng.integers(-50, max(2, n // 3), size=n).astype(np.int64)
  if n > 10:
    ids[3] = np.iinfo(np.int64).min  # the scratch sentinel value is a legal id too
    ids[7] = np.iinfo(np.int64).min
  u, idx = de.unique(torch.from_numpy(ids).cuda())
  eu, eidx = O.unique_first_occurrence(ids)
  np.testing.assert_array_equal(u.cpu().numpy(), eu)
  np.testing.assert_array_equal(idx.cpu().numpy(), eidx)


@pytest.mark.parametrize("S", [1, 2, 3, 8])
@pytest.mark.parametrize("gpu_mode", [True, False])
def test_partition_and_row_permutes(S, gpu_mode):
  torch = _torch()
  from recommenders_addons_b200.dynamic_embedding import variable as V
  rng = np.random.default_rng(S)
  n = 50000
  keys = rng.integers(-2**63, 2**63 - 1, size=n, dtype=np.int64)
  kt = torch.from_numpy(keys).cuda()
  grouped, perm, counts = V.partition(kt, S, gpu_mode)
  owner = O.default_partition_fn(keys, S, gpu_mode)
  order = np.argsort(owner, kind="stable")  # dynamic_partition keeps the original order inside a partition
  np.testing.assert_array_equal(perm.cpu().numpy(), order)
  np.testing.assert_array_equal(grouped.cpu().numpy(), keys[order])
  np.testing.assert_array_equal(counts.cpu().numpy(), np.bincount(owner, minlength=S))
  for dim in (1, 5, 64):
    rows = torch.randn(n, dim, device="cuda")
    g = V.gather_rows(rows, perm)
    assert torch.equal(g, rows[perm.long()])
    assert torch.equal(V.scatter_rows(g, perm), rows)


def _sparse_case(rng, batch, max_per_row, vocab, empty_rows=True):
  counts = rng.integers(0 if empty_rows else 1, max_per_row + 1, size=batch)
  counts[0] = max(1, counts[0])
  seg = np.repeat(np.arange(batch), counts).astype(np.int32)
  ids = rng.integers(0, vocab, size=seg.shape[0]).astype(np.int64)
  w = rng.uniform(0.25, 2.0, size=seg.shape[0]).astype(np.float32)
  return ids, seg, w


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("use_weights", [False, True])
@pytest.mark.parametrize("dim", [1, 5, 16, 64, 200])
def test_lookup_sparse_bit_exact_vs_oracle(combiner, use_weights, dim):
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  from recommenders_addons_b200.dynamic_embedding.ops import lookup_sparse_fused
  rng = np.random.default_rng(dim + 17)
  vocab, batch = 3000, 700
  var = de.Variable(dim=dim, initializer=0.25, name="sp-%d-%s-%d" % (dim, combiner, use_weights))
  dev = var.tables[0].device
  present = rng.choice(vocab, size=2000, replace=False).astype(np.int64)   # ~1/3 of ids are missing
  vals = rng.normal(0, 0.05, (2000, dim)).astype(np.float32)
  var.upsert(torch.from_numpy(present).to(dev), torch.from_numpy(vals).to(dev))
  ot = O.PortTable(dim)
  ot.insert(present, vals)
  ids, seg, w = _sparse_case(rng, batch, 9, vocab)
  if not use_weights:
    w = None
  got = lookup_sparse_fused(var, torch.from_numpy(ids), torch.from_numpy(seg),
                            None if w is None else torch.from_numpy(w), batch, combiner)
  exp = O.embedding_lookup_sparse(ot, ids, seg, w, batch, combiner, default=np.full(dim, 0.25, np.float32))
  np.testing.assert_array_equal(got.cpu().numpy(), exp)
  # public API (SparseTensor-shaped input) takes the same fused path
  ind = np.stack([seg.astype(np.int64), np.zeros_like(seg, dtype=np.int64)], 1)
  sp = de.SparseIds(torch.from_numpy(ind).to(dev), torch.from_numpy(ids).to(dev), (batch, 16))
  sw = None if w is None else de.SparseIds(sp.indices, torch.from_numpy(w).to(dev), (batch, 16))
  api = de.embedding_lookup_sparse(var, sp, sw, combiner=combiner)
  assert torch.equal(api, got)
  # composed path (unique -> lookup -> gather*w -> segment sum), reference tolerance
  comp, tw = de.embedding_lookup_sparse(var, sp, sw, combiner=combiner, return_trainable=True)
  np.testing.assert_allclose(comp.detach().cpu().numpy(), exp, rtol=1e-5, atol=1e-6)
  assert tw.ids.numel() == np.unique(ids).shape[0]


def test_lookup_sparse_c3_sized_bit_exact():
  """BASELINE configs[2] size: 26 features x batch 65536 = 1,703,936 ids, dim 64, Zipf-skewed with repeats, ~10 % of
  the ids absent -> det_lookup_sparse must equal the oracle BIT FOR BIT in both shapes the workload comes in:
  (A) one output row per (sample, feature) (bench.py --workload c3), (B) the 26 ids of a sample combined into one row
  (weighted mean).  The oracle's per-id Python loop is vectorised over ROWS here (26 sequential fp32 mul+add passes in
  id order, the same operations in the same order); that vectorisation is itself pinned against
  oracle.embedding_lookup_sparse on a prefix."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  from recommenders_addons_b200.dynamic_embedding.ops import lookup_sparse_fused
  rng = np.random.default_rng(2026)
  dim, nfeat, batch = 64, 26, 65536
  nnz = nfeat * batch
  vocab = np.maximum(1000, np.exp(rng.uniform(np.log(1e3), np.log(3e5), nfeat))).astype(np.int64)
  offs = np.concatenate([[0], np.cumsum(vocab)])[:-1]
  cols = []
  for f in range(nfeat):   # Zipf(1.05) over the feature's vocabulary, one id per (sample, feature)
    w = np.arange(1, vocab[f] + 1, dtype=np.float64) ** -1.05
    cdf = np.cumsum(w) / w.sum()
    cols.append(np.minimum(np.searchsorted(cdf, rng.random(batch)), vocab[f] - 1) + offs[f])
  ranks = np.stack(cols, 1).reshape(-1).astype(np.int64)                 # row-major (sample, feature)
  ids = (ranks * np.int64(2654435761) + 12345) ^ (ranks << 20)            # scrambled int64 keys, injective
  uniq = np.unique(ids)
  present = uniq[rng.random(uniq.shape[0]) < 0.9]
  vals = rng.normal(0, 0.05, (present.shape[0], dim)).astype(np.float32)
  var = de.Variable(dim=dim, initializer=0.125, init_size=4 * present.shape[0], name="sp-c3")
  dev = var.tables[0].device
  var.upsert(torch.from_numpy(present).to(dev), torch.from_numpy(vals).to(dev))
  ot = O.PortTable(dim)
  ot.insert(present, vals)
  default = np.full(dim, 0.125, np.float32)
  u, inv = O.unique_first_occurrence(ids)
  emb = ot.find(u, default)[inv]                                          # [nnz, dim] rows in id order
  t_ids = torch.from_numpy(ids).to(dev)
  # (A) one id per output row, combiner sum: the output is the gathered rows themselves (0 + row * 1)
  seg_a = torch.arange(nnz, dtype=torch.int32, device=dev)
  got_a = lookup_sparse_fused(var, t_ids, seg_a, None, nnz, "sum")
  exp_a = (np.zeros_like(emb) + emb * np.float32(1)).astype(np.float32)
  assert torch.equal(got_a.cpu(), torch.from_numpy(exp_a))
  del got_a
  # (B) 26 ids per row, weighted mean
  wts = rng.uniform(0.25, 2.0, nnz).astype(np.float32)
  seg_b = (np.arange(nnz) // nfeat).astype(np.int32)
  got_b = lookup_sparse_fused(var, t_ids, torch.from_numpy(seg_b).to(dev), torch.from_numpy(wts).to(dev), batch, "mean")
  e3, w3 = emb.reshape(batch, nfeat, dim), wts.reshape(batch, nfeat)
  out = np.zeros((batch, dim), np.float32)
  wsum = np.zeros(batch, np.float32)
  for j in range(nfeat):
    out = out + e3[:, j] * w3[:, j, None]
    wsum = wsum + w3[:, j]
  exp_b = (out / wsum[:, None]).astype(np.float32)
  np.testing.assert_array_equal(got_b.cpu().numpy(), exp_b)
  # the vectorised oracle == the per-id loop of oracle.embedding_lookup_sparse on a prefix of 1500 rows
  m = 1500 * nfeat
  ref = O.embedding_lookup_sparse(ot, ids[:m], seg_b[:m], wts[:m], 1500, "mean", default=default)
  np.testing.assert_array_equal(exp_b[:1500], ref)


def test_lookup_sparse_reference_test_vectors():
  """EmbeddingLookupSparseTest grouping (dynamic_embedding_ops_test.py:875-969), rtol=atol=1e-6 against the
  reference test's own NumPy formulation."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  from tests.test_oracle import _embedding_result
  for dim in (1, 5):
    rng = np.random.default_rng(3)
    vocab = 13
    params = {i: rng.standard_normal(dim).astype(np.float32) for i in range(vocab)}
    var = de.Variable(dim=dim, name="spref-%d" % dim)
    dev = var.tables[0].device
    var.upsert(torch.arange(vocab, device=dev), torch.from_numpy(np.stack([params[i] for i in range(vocab)])).to(dev))
    grouped_ids = [[0, 1, 2], [3], [4, 5, 6, 7, 7], [8, 9], [10, 11, 12, 0]]
    for use_w in (False, True):
      gw = [[rng.uniform(0.5, 2) for _ in g] for g in grouped_ids] if use_w else None
      ids = torch.tensor([i for g in grouped_ids for i in g], device=dev)
      ind = torch.tensor([[r, c] for r, g in enumerate(grouped_ids) for c, _ in enumerate(g)], device=dev)
      sp = de.SparseIds(ind, ids, (5, 5))
      sw = None if not use_w else de.SparseIds(ind, torch.tensor([x for g in gw for x in g], dtype=torch.float32,
                                                                  device=dev), (5, 5))
      v, ws, wsq = _embedding_result(params, grouped_ids, gw)
      for comb in ("sum", "mean", "sqrtn"):
        exp = v if comb == "sum" else (v / ws[:, None] if comb == "mean" else v / np.sqrt(wsq)[:, None])
        got = de.embedding_lookup_sparse(var, sp, sw, combiner=comb)
        np.testing.assert_allclose(got.cpu().numpy(), exp, rtol=1e-6, atol=1e-6)


def test_safe_embedding_lookup_sparse_fixture():
  """SafeEmbeddingLookupSparseTest fixture _ids_and_weights_2d (dynamic_embedding_ops_test.py:187-216):
  row0: ids 0,1 and invalid -100 (weights 1,2,1); row1: only -100; row2: empty; row3: id 2 (w 3);
  row4: ids 0,1 with weights 0 / -0.5 (pruned for mean)."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  dim = 4
  var = de.Variable(dim=dim, name="safe")
  dev = var.tables[0].device
  emb = torch.arange(3 * dim, dtype=torch.float32, device=dev).reshape(3, dim) + 1
  var.upsert(torch.tensor([0, 1, 2], device=dev), emb)
  ind = torch.tensor([[0, 0], [0, 1], [0, 2], [1, 0], [3, 0], [4, 0], [4, 1]], device=dev)
  ids = torch.tensor([0, 1, -100, -100, 2, 0, 1], device=dev)
  w = torch.tensor([1.0, 2.0, 1.0, 1.0, 3.0, 0.0, -0.5], device=dev)
  sp, sw = de.SparseIds(ind, ids, (5, dim)), de.SparseIds(ind, w, (5, dim))
  out = de.safe_embedding_lookup_sparse(var, sp, sw, combiner="mean")
  z = torch.zeros(dim, device=dev)
  exp = torch.stack([(emb[0] * 1 + emb[1] * 2 + z * 1) / 4.0, z, z, emb[2], z])
  torch.testing.assert_close(out, exp, rtol=1e-6, atol=1e-6)
  out = de.safe_embedding_lookup_sparse(var, sp, sw, combiner="mean", default_id=2)
  exp = torch.stack([(emb[0] * 1 + emb[1] * 2) / 4.0, z, emb[2], emb[2], emb[2]])
  torch.testing.assert_close(out, exp, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("dim", [1, 6, 16, 64, 128])
@pytest.mark.parametrize("eps", [0.0, 1e-7])
def test_fused_adagrad_twin(dim, eps):
  """Twin-model test (dynamic_embedding_optimizer_test.py:349-440): oracle tables stepped with
  find -> dense rule -> upsert  vs  the fused kernel, 10 steps, bit-exact params AND accumulators."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  rng = np.random.default_rng(dim)
  var = de.Variable(dim=dim, initializer=0.05, num_slot_planes=1, name="ada-%d-%g" % (dim, eps))
  dev = var.tables[0].device
  opt = de.DynamicEmbeddingOptimizer(de.FusedAdagrad(0.1, initial_accumulator_value=0.1, epsilon=eps))
  p, a = O.PortTable(dim), O.PortTable(dim)
  ip, ia = np.full(dim, 0.05, np.float32), np.full(dim, 0.1, np.float32)
  for step in range(10):
    keys = rng.choice(5000, size=1500, replace=False).astype(np.int64)
    g = rng.normal(0, 1e-2, (1500, dim)).astype(np.float32)
    O.sparse_adagrad_step(p, a, keys, g, 0.1, ip, ia, eps)
    opt.apply_gradients([(torch.from_numpy(g).to(dev), (var, torch.from_numpy(keys).to(dev)))])
  assert int(var.size()) == p.size()
  t = var.tables[0]
  for plane, ot in ((0, p), (1, a)):
    k, v = t.export(plane=plane)
    o = torch.argsort(k)
    ek, ev = sorted_export(ot)
    np.testing.assert_array_equal(k[o].cpu().numpy(), ek)
    np.testing.assert_array_equal(v[o].cpu().numpy(), ev)


@pytest.mark.parametrize("dim", [1, 6, 64])
def test_fused_adam_twin(dim):
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  rng = np.random.default_rng(dim + 100)
  var = de.Variable(dim=dim, initializer=0.0, num_slot_planes=2, name="adam-%d" % dim)
  dev = var.tables[0].device
  opt = de.FusedAdam(0.01, 0.9, 0.999, 1e-8)
  p, m, v = O.PortTable(dim), O.PortTable(dim), O.PortTable(dim)
  z = np.zeros(dim, np.float32)
  for step in range(1, 9):
    keys = rng.choice(3000, size=1000, replace=False).astype(np.int64)
    g = rng.normal(0, 1e-2, (1000, dim)).astype(np.float32)
    alpha = O.adam_scalars(0.01, 0.9, 0.999, step)
    O.sparse_adam_step(p, m, v, keys, g, alpha, 0.9, 0.999, 1e-8, z)
    opt.apply_gradients([(torch.from_numpy(g).to(dev), (var, torch.from_numpy(keys).to(dev)))])
    assert np.float32(opt.alpha()) == alpha
  t = var.tables[0]
  for plane, ot in ((0, p), (1, m), (2, v)):
    k, val = t.export(plane=plane)
    o = torch.argsort(k)
    ek, ev = sorted_export(ot)
    np.testing.assert_array_equal(k[o].cpu().numpy(), ek)
    np.testing.assert_array_equal(val[o].cpu().numpy(), ev)


def test_train_step_forward_backward_matches_dense_twin():
  """End-to-end: embedding_lookup_sparse(return_trainable) -> loss -> autograd -> fused Adagrad on the unique
  rows, against a dense torch.nn.Embedding-style twin trained with the same rule (tolerance 1e-6: the
  backward sum over duplicate ids is order-dependent in both frameworks)."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  dim, vocab, batch = 8, 50, 32
  torch.manual_seed(0)
  var = de.Variable(dim=dim, initializer=0.0, num_slot_planes=1, name="e2e-train")
  dev = var.tables[0].device
  opt = de.FusedAdagrad(0.1, 0.1)
  W = torch.zeros(vocab, dim, device=dev)
  A = torch.full((vocab, dim), 0.1, device=dev)
  target = torch.randn(batch, dim, device=dev)
  for step in range(5):
    ids = torch.randint(0, vocab, (batch * 3,), device=dev)
    ind = torch.stack([torch.arange(batch, device=dev).repeat_interleave(3), torch.arange(3, device=dev).repeat(batch)], 1)
    sp = de.SparseIds(ind, ids, (batch, 3))
    out, tw = de.embedding_lookup_sparse(var, sp, None, combiner="sum", return_trainable=True)
    loss = ((out - target) ** 2).sum()
    loss.backward()
    opt.apply_gradients([(tw.values.grad, tw)])
    Wd = W.clone().requires_grad_(True)
    outd = torch.zeros(batch, dim, device=dev).index_add(0, ind[:, 0], Wd[ids])
    ((outd - target) ** 2).sum().backward()
    touched = torch.unique(ids)
    g = Wd.grad[touched]
    A[touched] = A[touched] + g * g
    W[touched] = W[touched] - 0.1 * g / A[touched].sqrt()
  got = var.lookup(torch.arange(vocab, device=dev))
  torch.testing.assert_close(got, W, rtol=1e-5, atol=1e-6)


def test_slot_state_of_keys_created_by_insert():
  """A key written by insert/accum has NO optimizer slot yet (the reference keeps slots in separate tables,
  dynamic_embedding_optimizer.py:870-958): its first optimizer step must start from the slot initializer,
  and an export of the slot plane shows the initializer for never-stepped keys."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  dim = 16
  rng = np.random.default_rng(9)
  var = de.Variable(dim=dim, initializer=0.0, num_slot_planes=1, name="slot-lazy")
  dev = var.tables[0].device
  keys = np.arange(4000, dtype=np.int64)
  vals = rng.normal(0, 0.01, (4000, dim)).astype(np.float32)
  var.upsert(torch.from_numpy(keys).to(dev), torch.from_numpy(vals).to(dev))
  p, a = O.PortTable(dim), O.PortTable(dim)
  p.insert(keys, vals)
  opt = de.FusedAdagrad(0.1, initial_accumulator_value=0.1)
  step_keys = np.concatenate([keys[::3], np.arange(5000, 5500)]).astype(np.int64)  # stepped existing + brand new
  for _ in range(2):
    g = rng.normal(0, 1e-2, (step_keys.shape[0], dim)).astype(np.float32)
    O.sparse_adagrad_step(p, a, step_keys, g, 0.1, np.zeros(dim, np.float32), np.full(dim, 0.1, np.float32))
    opt.apply_gradients([(torch.from_numpy(g).to(dev), (var, torch.from_numpy(step_keys).to(dev)))])
  k, v = var.tables[0].export(plane=0)
  o = torch.argsort(k)
  ek, ev = sorted_export(p)
  np.testing.assert_array_equal(k[o].cpu().numpy(), ek)
  np.testing.assert_array_equal(v[o].cpu().numpy(), ev)
  k1, a1 = var.tables[0].export(plane=1)
  o1 = torch.argsort(k1)
  k1, a1 = k1[o1].cpu().numpy(), a1[o1].cpu().numpy()
  np.testing.assert_array_equal(k1, ek)
  stepped = np.isin(k1, step_keys)
  ak, av = sorted_export(a)
  np.testing.assert_array_equal(k1[stepped], ak)
  np.testing.assert_array_equal(a1[stepped], av)
  assert (a1[~stepped] == np.float32(0.1)).all()


def test_embedding_layer_twin_training():
  """A de.layers.SquashedEmbedding trained with the fused Adagrad == a dense torch twin trained with the same rule
  (the comparison the reference makes between HvdAllToAllEmbedding and tf.keras.layers.Embedding,
  kernel_tests/horovod_sync_train_test.py:300-336)."""
  torch = _torch()
  from recommenders_addons_b200 import dynamic_embedding as de
  dim, vocab, batch, n = 8, 40, 16, 3
  torch.manual_seed(1)
  layer = de.layers.SquashedEmbedding(dim, combiner="sum", initializer=0.0, name="layer-twin", num_slot_planes=1)
  layer.train()
  opt = de.FusedAdagrad(0.1, 0.1)
  dev = layer.params.tables[0].device
  W = torch.zeros(vocab, dim, device=dev)
  A = torch.full((vocab, dim), 0.1, device=dev)
  target = torch.randn(batch, dim, device=dev)
  for step in range(6):
    ids = torch.randint(0, vocab, (batch, n), device=dev)
    out = layer(ids)
    assert out.shape == (batch, dim)
    ((out - target) ** 2).sum().backward()
    layer.apply_gradients(opt)
    Wd = W.clone().requires_grad_(True)
    ((Wd[ids].sum(1) - target) ** 2).sum().backward()
    touched = torch.unique(ids)
    g = Wd.grad[touched]
    A[touched] = A[touched] + g * g
    W[touched] = W[touched] - 0.1 * g / A[touched].sqrt()
  got = layer.params.lookup(torch.arange(vocab, device=dev))
  torch.testing.assert_close(got, W, rtol=1e-5, atol=1e-6)
  plain = de.layers.Embedding(dim, name="layer-plain")
  assert plain(torch.tensor([[15, 2], [4, 92], [22, 4]], device=dev)).shape == (3, 2, dim)
