"""fused.cu on the SIMT emulator (tests/emu/), through the real C ABI: det_unique, det_partition / gather / scatter,
det_lookup_sparse and the fused Adagrad / Adam steps, compared BIT-EXACTLY with the NumPy restatements of
oracle/oracle.py -- the same checks tests/test_fused_gpu.py makes on the B200, at sizes the emulator runs in seconds.
(The emulated build uses the plain-load tile schedule where the GPU build stages key tiles by TMA; the arithmetic,
the probe / claim protocol and the host code are the same source.)"""

import numpy as np
import pytest

from oracle import oracle as O
from recommenders_addons_b200 import _lib as real
from tests.helpers import sorted_export
from tests.test_detable_emu import L, P, Table, ck

_FUSED = ["det_apply_dup_workspace_bytes", "det_apply_adagrad_dup", "det_apply_adam_dup",
          "det_unique_workspace_bytes", "det_unique", "det_lookup_sparse", "det_lookup_sparse_clip", "det_apply_adagrad",
          "det_apply_adam",
          "det_partition_workspace_bytes", "det_partition", "det_scatter_rows", "det_gather_rows"]


def F():
  l = L()
  if not getattr(l, "_fused_ready", False):
    for name in _FUSED:
      res, args = real.SIGNATURES[name]
      fn = getattr(l, name)
      fn.restype, fn.argtypes = res, args
    l._fused_ready = True
  return l


@pytest.mark.parametrize("n", [1, 31, 1000, 5000])
def test_unique_first_occurrence(n):
  rng = np.random.default_rng(n)
  ids = rng.integers(-50, max(2, n // 3), size=n).astype(np.int64)
  if n > 10:
    ids[3] = ids[7] = np.iinfo(np.int64).min
  u = np.empty(n, dtype=np.int64)
  idx = np.empty(n, dtype=np.int32)
  cnt = np.zeros(1, dtype=np.int64)
  wsb = F().det_unique_workspace_bytes(n)
  ws = np.empty(wsb, dtype=np.uint8)
  ck(F().det_unique(P(ids), n, P(u), P(idx), P(cnt), P(ws), wsb, None))
  eu, eidx = O.unique_first_occurrence(ids)
  np.testing.assert_array_equal(u[:cnt[0]], eu)
  np.testing.assert_array_equal(idx, eidx)


@pytest.mark.parametrize("S,gpu_mode", [(1, True), (2, True), (3, False), (8, True)])
def test_partition_and_row_permutes(S, gpu_mode):
  rng = np.random.default_rng(S)
  n = 3000
  keys = rng.integers(-2**63, 2**63 - 1, size=n, dtype=np.int64)
  grouped = np.empty(n, dtype=np.int64)
  perm = np.empty(n, dtype=np.int32)
  counts = np.zeros(S, dtype=np.int64)
  wsb = F().det_partition_workspace_bytes(n, S)
  ws = np.empty(wsb, dtype=np.uint8)
  ck(F().det_partition(P(keys), n, S, int(gpu_mode), P(grouped), P(perm), P(counts), P(ws), wsb, None))
  owner = O.default_partition_fn(keys, S, gpu_mode)
  order = np.argsort(owner, kind="stable")
  np.testing.assert_array_equal(perm, order)
  np.testing.assert_array_equal(grouped, keys[order])
  np.testing.assert_array_equal(counts, np.bincount(owner, minlength=S))
  for dim in (1, 5, 16):
    rows = rng.standard_normal((n, dim)).astype(np.float32)
    g = np.empty_like(rows)
    ck(F().det_gather_rows(P(rows), P(perm), n, dim * 4, P(g), None))
    np.testing.assert_array_equal(g, rows[perm])
    back = np.empty_like(rows)
    ck(F().det_scatter_rows(P(g), P(perm), n, dim * 4, P(back), None))
    np.testing.assert_array_equal(back, rows)


def _sparse_case(rng, batch, max_per_row, vocab):
  counts = rng.integers(0, max_per_row + 1, size=batch)
  counts[0] = max(1, counts[0])
  seg = np.repeat(np.arange(batch), counts).astype(np.int32)
  ids = rng.integers(0, vocab, size=seg.shape[0]).astype(np.int64)
  w = rng.uniform(0.25, 2.0, size=seg.shape[0]).astype(np.float32)
  return ids, seg, w


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("use_weights", [False, True])
@pytest.mark.parametrize("dim", [1, 5, 16, 200])
def test_lookup_sparse_bit_exact_vs_oracle(combiner, use_weights, dim):
  rng = np.random.default_rng(dim + 17)
  vocab, batch = 600, 150
  t = Table(dim=dim, init=2048)
  present = rng.choice(vocab, size=400, replace=False).astype(np.int64)
  vals = rng.normal(0, 0.05, (400, dim)).astype(np.float32)
  t.insert(present, vals)
  ot = O.PortTable(dim)
  ot.insert(present, vals)
  ids, seg, w = _sparse_case(rng, batch, 9, vocab)
  if not use_weights:
    w = None
  default = np.full(dim, 0.25, np.float32)
  out = np.empty((batch, dim), dtype=np.float32)
  ck(F().det_lookup_sparse(t.h, P(ids), P(seg), P(w), len(ids), batch, real.COMBINERS[combiner], P(default), P(out), None))
  exp = O.embedding_lookup_sparse(ot, ids, seg, w, batch, combiner, default=default)
  np.testing.assert_array_equal(out, exp)
  t.close()


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("dim", [1, 6, 16, 64, 200])
@pytest.mark.parametrize("shape", ["identity", "same_count_not_identity", "identity_with_weights"])
def test_lookup_sparse_one_id_per_row(combiner, dim, shape):
  """nnz == batch: the library launches the find-shaped kernel AND the general pair and lets a device flag (are the
  segment ids 0..nnz-1?) pick the one that works.  identity: one id per row (the Criteo shape) -> rows / default rows,
  whatever the combiner; same_count_not_identity: nnz == batch but two ids share a row and one row is empty -> the
  general kernels must do the work; with weights the fast path is never tried.  All bit-exact vs the oracle."""
  rng = np.random.default_rng(dim * 5 + len(shape))
  vocab, batch = 500, 333
  t = Table(dim=dim, init=2048)
  present = rng.choice(vocab, size=350, replace=False).astype(np.int64)
  vals = rng.normal(0, 0.05, (350, dim)).astype(np.float32)
  t.insert(present, vals)
  ot = O.PortTable(dim)
  ot.insert(present, vals)
  ids = rng.integers(0, vocab, size=batch).astype(np.int64)           # repeats and ~30 % missing ids
  seg = np.arange(batch, dtype=np.int32)
  w = None
  if shape == "same_count_not_identity":
    seg[100] = 99                                                      # rows 99 (two ids) and 100 (none)
  if shape == "identity_with_weights":
    w = rng.uniform(0.25, 2.0, size=batch).astype(np.float32)
  default = np.full(dim, 0.25, np.float32)
  out = np.full((batch, dim), np.nan, dtype=np.float32)
  ck(F().det_lookup_sparse(t.h, P(ids), P(seg), P(w), batch, batch, real.COMBINERS[combiner], P(default), P(out), None))
  exp = O.embedding_lookup_sparse(ot, ids, seg, w, batch, combiner, default=default)
  np.testing.assert_array_equal(out, exp)
  t.close()


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("dim", [1, 5, 16, 64, 128])
def test_lookup_sparse_with_max_norm(combiner, dim):
  """det_lookup_sparse_clip: tf.clip_by_norm of every looked-up row (missing ids: of the default row) BEFORE the weighted
  combine (embedding_lookup_sparse(..., max_norm), python/ops/embedding_weights.py:497-521).  The norm is a sum of dim
  squares (summation order differs from NumPy's): 1e-6 like the reference's max_norm tests, not bit-exact."""
  rng = np.random.default_rng(dim + 31)
  vocab, batch, max_norm = 600, 150, 0.35
  t = Table(dim=dim, init=2048)
  present = rng.choice(vocab, size=400, replace=False).astype(np.int64)
  vals = rng.normal(0, 0.05 * 8 / np.sqrt(dim + 7), (400, dim)).astype(np.float32)   # norms on both sides of max_norm
  t.insert(present, vals)
  ids, seg, w = _sparse_case(rng, batch, 9, vocab)
  default = np.full(dim, 0.9 / np.sqrt(dim), np.float32)                 # ||default|| = 0.9 > max_norm: clipped too
  out = np.empty((batch, dim), dtype=np.float32)
  ck(F().det_lookup_sparse_clip(t.h, P(ids), P(seg), P(w), len(ids), batch, real.COMBINERS[combiner], P(default), max_norm,
                                P(out), None))

  def clip(x):
    n = np.sqrt((x.astype(np.float64) ** 2).sum(-1, keepdims=True))
    return (x * (max_norm / np.maximum(n, max_norm))).astype(np.float32)
  ot = O.PortTable(dim)
  ot.insert(present, clip(vals))
  exp = O.embedding_lookup_sparse(ot, ids, seg, w, batch, combiner, default=clip(default[None])[0])
  norms = np.sqrt((vals.astype(np.float64) ** 2).sum(-1))
  assert (norms > max_norm).any() and (norms < max_norm).any()
  np.testing.assert_allclose(out, exp, rtol=1e-6, atol=1e-6)
  # max_norm = 0 is the unclipped kernel, bit for bit
  out0, ref0 = np.empty_like(out), np.empty_like(out)
  ck(F().det_lookup_sparse_clip(t.h, P(ids), P(seg), P(w), len(ids), batch, real.COMBINERS[combiner], P(default), 0.0, P(out0), None))
  ck(F().det_lookup_sparse(t.h, P(ids), P(seg), P(w), len(ids), batch, real.COMBINERS[combiner], P(default), P(ref0), None))
  np.testing.assert_array_equal(out0, ref0)
  assert F().det_lookup_sparse_clip(t.h, P(ids), P(seg), P(w), len(ids), batch, 0, P(default), -1.0, P(out), None) == 1
  t.close()


def test_lookup_sparse_max_norm_wide_rows_are_refused():
  t = Table(dim=200, init=256)
  ids, seg = np.arange(4, dtype=np.int64), np.arange(4, dtype=np.int32)
  out = np.empty((4, 200), dtype=np.float32)
  dflt = np.zeros(200, np.float32)
  st = F().det_lookup_sparse_clip(t.h, P(ids), P(seg), None, 4, 4, 0, P(dflt), 1.0, P(out), None)
  assert st == 5 and b"composed" in L().det_last_error()          # DET_UNIMPLEMENTED: the Python mirror composes
  t.close()


def _export_sorted(t, plane):
  k, v = t.export(plane)
  o = np.argsort(k)
  return k[o], v[o]


@pytest.mark.parametrize("dim", [1, 6, 16, 64])
@pytest.mark.parametrize("eps", [0.0, 1e-7])
def test_fused_adagrad_twin(dim, eps):
  """dynamic_embedding_optimizer_test.py:349-440 shape: oracle tables stepped find -> dense rule -> upsert vs the fused
  kernel, bit-exact params AND accumulators; the table grows underneath"""
  rng = np.random.default_rng(dim)
  t = Table(dim=dim, init=256, slot_planes=1)
  p, a = O.PortTable(dim), O.PortTable(dim)
  ip, ia = np.full(dim, 0.05, np.float32), np.full(dim, 0.1, np.float32)
  for step in range(6):
    keys = rng.choice(1200, size=400, replace=False).astype(np.int64)
    g = rng.normal(0, 1e-2, (400, dim)).astype(np.float32)
    O.sparse_adagrad_step(p, a, keys, g, 0.1, ip, ia, eps)
    ck(F().det_apply_adagrad(t.h, P(keys), P(g), 400, 0.1, eps, P(ip), 0, 0.1, None))
  assert t.size() == p.size()
  for plane, ot in ((0, p), (1, a)):
    k, v = _export_sorted(t, plane)
    ek, ev = sorted_export(ot)
    np.testing.assert_array_equal(k, ek)
    np.testing.assert_array_equal(v, ev)
  t.close()


@pytest.mark.parametrize("dim", [1, 6, 64])
def test_fused_adam_twin(dim):
  rng = np.random.default_rng(dim + 100)
  t = Table(dim=dim, init=256, slot_planes=2)
  p, m, v = O.PortTable(dim), O.PortTable(dim), O.PortTable(dim)
  z = np.zeros(dim, np.float32)
  for step in range(1, 6):
    keys = rng.choice(900, size=300, replace=False).astype(np.int64)
    g = rng.normal(0, 1e-2, (300, dim)).astype(np.float32)
    alpha = O.adam_scalars(0.01, 0.9, 0.999, step)
    O.sparse_adam_step(p, m, v, keys, g, alpha, 0.9, 0.999, 1e-8, z)
    ck(F().det_apply_adam(t.h, P(keys), P(g), 300, float(alpha), 0.9, 0.999, 1e-8, P(z), 0, None))
  for plane, ot in ((0, p), (1, m), (2, v)):
    k, val = _export_sorted(t, plane)
    ek, ev = sorted_export(ot)
    np.testing.assert_array_equal(k, ek)
    np.testing.assert_array_equal(val, ev)
  t.close()


@pytest.mark.parametrize("dim,opt", [(64, "adagrad"), (16, "adagrad"), (32, "adam")])
def test_apply_with_repeated_ids_in_one_call(dim, opt):
  """det_apply_*_dup: ids WITH repeats (Zipf head: one id owns a group of > 1024 rows, the column-sliced path) and their
  row gradients in ONE call -- unique + position-order sum + fused step chained on the device, the unique count never
  read by the host.  Twin: oracle unique_first_occurrence -> segment_reduce -> sparse_*_step; params and slots bit-exact
  over several steps, the table growing underneath; the device count is returned too."""
  rng = np.random.default_rng(dim * 3 + len(opt))
  planes = 1 if opt == "adagrad" else 2
  t = Table(dim=dim, init=256, slot_planes=planes)
  tabs = [O.PortTable(dim) for _ in range(1 + planes)]
  ip = np.full(dim, 0.05, np.float32)
  ia = np.full(dim, 0.1, np.float32)
  z = np.zeros(dim, np.float32)
  n = 2600
  wsb = F().det_apply_dup_workspace_bytes(n, dim)
  raw = np.zeros(wsb + 256, np.uint8)
  ws = raw[(-raw.ctypes.data) % 256:][:wsb]
  cnt = np.zeros(1, np.int64)
  for step in range(1, 4):
    ids = np.concatenate([np.full(1200, 7), rng.integers(0, 40, 600), rng.integers(0, 5000, n - 1800)]).astype(np.int64)
    rng.shuffle(ids)
    g = (rng.normal(0, 1e-2, (n, dim)) * np.exp(rng.uniform(-4, 4, (n, 1)))).astype(np.float32)
    uniq, idx = O.unique_first_occurrence(ids)
    gs = O.segment_reduce(g, idx, len(uniq))
    if opt == "adagrad":
      O.sparse_adagrad_step(tabs[0], tabs[1], uniq, gs, 0.1, ip, ia, 0.0)
      ck(F().det_apply_adagrad_dup(t.h, P(ids), P(g), n, 0.1, 0.0, P(ip), 0.1, P(ws), wsb, P(cnt), None))
    else:
      alpha = O.adam_scalars(0.01, 0.9, 0.999, step)
      O.sparse_adam_step(tabs[0], tabs[1], tabs[2], uniq, gs, alpha, 0.9, 0.999, 1e-8, ip)
      ck(F().det_apply_adam_dup(t.h, P(ids), P(g), n, float(alpha), 0.9, 0.999, 1e-8, P(ip), P(ws), wsb, P(cnt), None))
    assert cnt[0] == len(uniq)
  assert t.size() == tabs[0].size()
  for plane, ot in enumerate(tabs):
    k, v = _export_sorted(t, plane)
    ek, ev = sorted_export(ot)
    np.testing.assert_array_equal(k, ek)
    np.testing.assert_array_equal(v, ev)
  # argument checks: workspace too small / misaligned, odd rows are refused with DET_UNIMPLEMENTED (5)
  assert F().det_apply_adagrad_dup(t.h, P(ids), P(g), n, 0.1, 0.0, P(ip), 0.1, P(ws), wsb - 1024, None, None) == 1
  t.close()
  t3 = Table(dim=3, init=256, slot_planes=2)
  g3 = np.zeros((8, 3), np.float32)
  i3 = np.arange(8, dtype=np.int64)
  w3b = F().det_apply_dup_workspace_bytes(8, 3)
  raw3 = np.zeros(w3b + 256, np.uint8)
  w3 = raw3[(-raw3.ctypes.data) % 256:][:w3b]
  assert F().det_apply_adagrad_dup(t3.h, P(i3), P(g3), 8, 0.1, 0.0, P(np.zeros(3, np.float32)), 0.1, P(w3), w3b, None, None) == 5
  t3.close()


def test_slot_state_of_keys_created_by_insert():
  """a key written by insert has NO optimizer slot yet: its first step starts from the slot initializer, and the slot
  plane exports the initializer for never-stepped keys (dynamic_embedding_optimizer.py:870-958)"""
  dim = 16
  rng = np.random.default_rng(9)
  t = Table(dim=dim, init=4096, slot_planes=1)
  keys = np.arange(800, dtype=np.int64)
  vals = rng.normal(0, 0.01, (800, dim)).astype(np.float32)
  t.insert(keys, vals)
  p, a = O.PortTable(dim), O.PortTable(dim)
  p.insert(keys, vals)
  step_keys = np.concatenate([keys[::3], np.arange(1000, 1100)]).astype(np.int64)
  z, ia = np.zeros(dim, np.float32), np.full(dim, 0.1, np.float32)
  for _ in range(2):
    g = rng.normal(0, 1e-2, (len(step_keys), dim)).astype(np.float32)
    O.sparse_adagrad_step(p, a, step_keys, g, 0.1, z, ia)
    ck(F().det_apply_adagrad(t.h, P(step_keys), P(g), len(step_keys), 0.1, 0.0, P(z), 0, 0.1, None))
  k, v = _export_sorted(t, 0)
  ek, ev = sorted_export(p)
  np.testing.assert_array_equal(k, ek)
  np.testing.assert_array_equal(v, ev)
  k1, a1 = _export_sorted(t, 1)
  stepped = np.isin(k1, step_keys)
  ak, av = sorted_export(a)
  np.testing.assert_array_equal(k1[stepped], ak)
  np.testing.assert_array_equal(a1[stepped], av)
  assert (a1[~stepped] == np.float32(0.1)).all()
  t.close()


def test_fused_optimizer_on_a_bounded_table_refreshes_scores():
  """det_apply_adagrad on a table with an eviction strategy: room is made by eviction, the stepped keys are touched"""
  dim = 8
  t = Table(dim=dim, init=512, max_capacity=512, strategy=1, slot_planes=1)      # LFU
  z = np.zeros(dim, np.float32)
  hot = np.arange(20, dtype=np.int64)
  nxt = 1000
  for step in range(8):
    keys = np.concatenate([hot, np.arange(nxt, nxt + 100, dtype=np.int64)])
    nxt += 100
    g = np.full((len(keys), dim), 0.01, np.float32)
    ck(F().det_apply_adagrad(t.h, P(keys), P(g), len(keys), 0.1, 0.0, P(z), 0, 0.1, None))
  assert t.size() <= 512 and t.stats()["evict_events"] >= 1
  assert (t.scores_of(hot) == 8).all() and t.find(hot)[1].all()       # the hot keys were touched every step and survive
  t.check()
  t.close()
