"""`de.segment_reduce` / det_segment_reduce (csrc/fused.cu K9): the per-unique-key gradient sum of the sparse optimizer
path (TF's _deduplicate_indexed_slices = unsorted_segment_sum before _resource_apply_sparse_duplicate_indices,
python/ops/dynamic_embedding_optimizer.py:150,184) -- rows of one key added in POSITION ORDER, so the result is
bit-identical to the sequential CPU sum (oracle.segment_reduce = np.add.at) and independent of the schedule.

First hardware run: round 1's driver box (all five suites passed on a fresh B200); ungated in round 2."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"     # tests/test_segreduce_emu.py re-runs these bodies over the emulated library with DEV = "cpu"
SCALE = 1        # the emulator runs the large cases at 1/64 of the size


def _de():
  from recommenders_addons_b200 import dynamic_embedding as de
  return de


def _rows(rng, n, dim):
  return (rng.normal(0, 1, (n, dim)) * np.exp(rng.uniform(-8, 8, (n, 1)))).astype(np.float32)


def _check(rows, idx, n_groups):
  got = _de().segment_reduce(torch.as_tensor(rows, device=DEV), torch.as_tensor(idx, device=DEV), n_groups)
  np.testing.assert_array_equal(got.cpu().numpy(), O.segment_reduce(rows, idx, n_groups))


@pytest.mark.parametrize("n,n_groups,dim", [(1, 1, 4), (1000, 37, 16), (50000, 9000, 64), (50000, 70000, 128), (30011, 257, 12),
                                            (4000, 50, 260)])
def test_segment_reduce_random_bit_exact(n, n_groups, dim):
  rng = np.random.default_rng(n + dim)
  _check(_rows(rng, n, dim), rng.integers(0, n_groups, size=n).astype(np.int32), n_groups)


def test_segment_reduce_criteo_shaped_step():
  """BASELINE configs[2] shape: 26 features x batch 65536 ids, Zipf(1.05) per feature -> det_unique -> per-unique sum;
  the head keys have thousands of rows each (the CTA-cooperative path), most keys one or two"""
  de = _de()
  rng = np.random.default_rng(45)
  nfeat, batch, dim = 26, 65536 // SCALE, 64
  vocab = np.maximum(1000, np.exp(rng.uniform(np.log(1e3), np.log(4e7), nfeat)) / SCALE).astype(np.int64)
  cols = [np.minimum(rng.zipf(1.05, size=batch), v) + o for v, o in zip(vocab, np.cumsum(vocab) - vocab)]
  ids = np.stack(cols, 1).reshape(-1).astype(np.int64)
  uniq, idx = de.unique(torch.as_tensor(ids, device=DEV))
  eu, eidx = O.unique_first_occurrence(ids)
  assert np.array_equal(uniq.cpu().numpy(), eu) and np.array_equal(idx.cpu().numpy(), eidx)
  assert np.bincount(eidx).max() > 64
  g = (rng.normal(0, 1e-2, (ids.shape[0], dim))).astype(np.float32)
  got = de.segment_reduce(torch.as_tensor(g, device=DEV), idx, uniq.numel())
  np.testing.assert_array_equal(got.cpu().numpy(), O.segment_reduce(g, eidx, eu.shape[0]))
  # and it is what the schedule-ordered atomics compute, up to fp32 rounding
  ref = torch.zeros_like(got).index_add_(0, idx.long(), torch.as_tensor(g, device=DEV))
  torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-3)   # sums of up to ~40 K rows in a different order


def test_segment_reduce_is_deterministic():
  rng = np.random.default_rng(3)
  n, n_groups, dim = 200000 // SCALE, 5000 // SCALE + 1, 64
  rows = torch.as_tensor(_rows(rng, n, dim), device=DEV)
  idx = torch.as_tensor(np.minimum(rng.zipf(1.1, size=n) - 1, n_groups - 1).astype(np.int32), device=DEV)
  a = _de().segment_reduce(rows, idx, n_groups)
  for _ in range(3):
    assert torch.equal(a, _de().segment_reduce(rows, idx, n_groups))


def test_segment_reduce_edge_cases():
  de = _de()
  rng = np.random.default_rng(8)
  # empty input, empty groups, dropped indices
  out = de.segment_reduce(torch.zeros((0, 8), device=DEV), torch.zeros(0, dtype=torch.int32, device=DEV), 5)
  assert out.shape == (5, 8) and not out.any()
  rows = _rows(rng, 700, 8)
  idx = rng.integers(-2, 12, size=700).astype(np.int32)
  _check(rows, idx, 10)
  with pytest.raises(TypeError):
    de.segment_reduce(torch.zeros((4, 8), dtype=torch.float64, device=DEV), torch.zeros(4, dtype=torch.int32, device=DEV), 2)
  with pytest.raises(ValueError):
    de.segment_reduce(torch.zeros((3, 8), device=DEV), torch.zeros(4, dtype=torch.int32, device=DEV), 2)


def test_sharded_combine_with_the_deterministic_reduce(monkeypatch):
  """PeerShardedVariable.apply_gradients with DET_GRAD_REDUCE=det: same parameters as the index_add combine (1e-6)
  and bit-identical to the oracle's sequential step"""
  de = _de()
  dim, world = 16, 1    # one shard: apply_gradients steps the local rank's inbox, so every routed key is updated
  rng = np.random.default_rng(21)
  keys = rng.choice(np.arange(1, 5000), 600, replace=False).astype(np.int64)
  dup = rng.choice(keys, 1500).astype(np.int64)                       # the same row from several "ranks"
  g = rng.normal(0, 1e-2, (dup.shape[0], dim)).astype(np.float32)
  results = []
  for mode in ("det", "torch"):
    monkeypatch.setenv("DET_GRAD_REDUCE", mode)
    shards = [de.Variable(dim=dim, init_size=1 << 13, initializer=0.25, num_slot_planes=1, devices=[DEV],
                          name="segred-%s-%d" % (mode, i)) for i in range(world)]
    pv = de.PeerShardedVariable(fake_shards=shards)
    pv.attach_inbox(4096)
    opt = de.FusedAdagrad(0.05, 0.1)
    pv.apply_gradients(opt, torch.as_tensor(dup, device=DEV), torch.as_tensor(g, device=DEV))
    results.append(pv.lookup(torch.as_tensor(keys, device=DEV)).cpu().numpy())
    pv.close()
  np.testing.assert_allclose(results[0], results[1], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("mode", ["det", "torch"])
def test_optimizer_step_with_duplicate_ids(monkeypatch, mode):
  """apply_sparse_duplicate_indices = unique + per-id gradient sum + ONE fused step (the reference's
  _resource_apply_sparse_duplicate_indices).  With DET_GRAD_REDUCE=det every step is bit-identical to the oracle
  (sequential sums, then the Adagrad rule); the index_add combine agrees to fp32 rounding."""
  monkeypatch.setenv("DET_GRAD_REDUCE", mode)
  de = _de()
  dim = 16
  rng = np.random.default_rng(33)
  var = de.Variable(dim=dim, init_size=1 << 13, initializer=0.5, num_slot_planes=1, devices=[DEV], name="dup-ids-" + mode)
  opt = de.FusedAdagrad(0.05, 0.1)
  p, a = O.PortTable(dim), O.PortTable(dim)
  for step in range(3):
    ids = np.minimum(rng.zipf(1.2, size=3000), 700).astype(np.int64) * 31 - 9
    g = rng.normal(0, 1e-2, (ids.shape[0], dim)).astype(np.float32)
    opt.iterations += 1
    opt.apply_sparse_duplicate_indices(var, torch.as_tensor(ids, device=DEV), torch.as_tensor(g, device=DEV))
    eu, eidx = O.unique_first_occurrence(ids)
    O.sparse_adagrad_step(p, a, eu, O.segment_reduce(g, eidx, eu.shape[0]), 0.05, np.full(dim, 0.5, np.float32),
                          np.full(dim, 0.1, np.float32))
  k, v = var.export()
  o = torch.argsort(k)
  ek, ev = p.export()
  eo = np.argsort(ek)
  np.testing.assert_array_equal(k[o].cpu().numpy(), ek[eo])
  if mode == "det":
    np.testing.assert_array_equal(v[o].cpu().numpy(), ev[eo])
  else:
    np.testing.assert_allclose(v[o].cpu().numpy(), ev[eo], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("mode", ["det", "torch"])
def test_embedding_lookup_unique_backward_is_the_gradient_dedupe(monkeypatch, mode):
  """d(embedding_lookup_unique)/d(unique rows) = per-unique sum of the output gradients (the gradient of the reference's
  gather, summed by _deduplicate_indexed_slices); bit-identical to the sequential sum under DET_GRAD_REDUCE=det"""
  monkeypatch.setenv("DET_GRAD_REDUCE", mode)
  de = _de()
  dim = 8
  rng = np.random.default_rng(41)
  var = de.Variable(dim=dim, init_size=1 << 12, initializer=0.1, devices=[DEV], name="lookup-unique-bwd-" + mode)
  ids = np.minimum(rng.zipf(1.3, size=2000), 300).astype(np.int64)
  out, tw = de.embedding_lookup_unique(var, torch.as_tensor(ids, device=DEV), return_trainable=True)
  gout = rng.normal(0, 1, (ids.shape[0], dim)).astype(np.float32)
  out.backward(torch.as_tensor(gout, device=DEV))
  eu, eidx = O.unique_first_occurrence(ids)
  exp = O.segment_reduce(gout, eidx, eu.shape[0])
  got = tw.values.grad.cpu().numpy()
  if mode == "det":
    np.testing.assert_array_equal(got, exp)
  else:
    np.testing.assert_allclose(got, exp, rtol=1e-5, atol=1e-3)   # atomics: sums of up to ~700 N(0,1) rows in any order


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
def test_trainable_sparse_lookup_through_the_fused_segment_sum(monkeypatch, combiner):
  """DET_SPARSE_TRAIN_FUSED=1: embedding_lookup_sparse(return_trainable=True) sums the TrainableWrapper's rows with
  det_sparse_segment_sum -- same values as the forward-only fused kernel, and the same parameters after one Adagrad
  step as the torch restatement of the path (1e-6)"""
  de = _de()
  dim, vocab, batch = 16, 200, 64
  rng = np.random.default_rng(50 + len(combiner))
  res = []
  for flag in ("1", "0"):
    monkeypatch.setenv("DET_SPARSE_TRAIN_FUSED", flag)
    var = de.Variable(dim=dim, init_size=1 << 12, initializer=0.0, num_slot_planes=1, devices=[DEV],
                      name="sparse-train-%s-%s" % (combiner, flag))
    var.upsert(torch.arange(vocab, device=DEV), torch.as_tensor(np.random.default_rng(1).normal(0, 0.1, (vocab, dim)).astype(np.float32), device=DEV))
    r = np.random.default_rng(7)
    ids = torch.as_tensor(r.integers(0, vocab, batch * 3), device=DEV)
    ind = torch.stack([torch.arange(batch, device=DEV).repeat_interleave(3), torch.arange(3, device=DEV).repeat(batch)], 1)
    w = torch.as_tensor(r.uniform(0.5, 2, batch * 3).astype(np.float32), device=DEV)
    sp, sw = de.SparseIds(ind, ids, (batch, 3)), de.SparseIds(ind, w, (batch, 3))
    out, tw = de.embedding_lookup_sparse(var, sp, sw, combiner=combiner, return_trainable=True)
    if flag == "1":
      assert torch.equal(out.detach(), de.embedding_lookup_sparse(var, sp, sw, combiner=combiner))
    (out * out).sum().backward()
    opt = de.FusedAdagrad(0.1, 0.1)
    opt.apply_gradients([(tw.values.grad, tw)])
    res.append(var.lookup(torch.arange(vocab, device=DEV)).cpu().numpy())
  np.testing.assert_allclose(res[0], res[1], rtol=1e-5, atol=1e-6)
