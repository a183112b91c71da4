"""det_segment_reduce (csrc/fused.cu K9: stable radix grouping + in-order row sums) on the SIMT emulator through the
real C ABI, compared BIT-EXACTLY with oracle.segment_reduce (np.add.at = the position order of TF's CPU
unsorted_segment_sum).  Covers: short groups, long groups (the CTA-cooperative path, > 64 rows of one index), every
radix-pass count (n_groups < 2^8, < 2^16, >= 2^16), odd dims (scalar path), wide rows (several vectors per lane, and
more columns than one staged tile), empty groups, dropped (negative / too large) indices, n = 0."""
import numpy as np
import pytest

from oracle import oracle as O
from recommenders_addons_b200 import _lib as real
from tests.test_detable_emu import L, P, ck

_FUNCS = ["det_segment_reduce_workspace_bytes", "det_segment_reduce", "det_unique_workspace_bytes", "det_unique"]


def S():
  l = L()
  if not getattr(l, "_segreduce_ready", False):
    for name in _FUNCS:
      res, args = real.SIGNATURES[name]
      fn = getattr(l, name)
      fn.restype, fn.argtypes = res, args
    l._segreduce_ready = True
  return l


def reduce_emu(rows, idx, n_groups):
  n, dim = rows.shape
  out = np.full((n_groups, dim), np.float32(np.nan))   # every row must be written
  wsb = S().det_segment_reduce_workspace_bytes(n, n_groups)
  ws = np.empty(wsb, dtype=np.uint8)
  ck(S().det_segment_reduce(P(rows), P(idx), n, n_groups, dim, P(out), P(ws), wsb, None))
  return out


def rows_of(rng, n, dim):
  # wide dynamic range: a different summation order would show up in the low bits
  return (rng.normal(0, 1, (n, dim)) * np.exp(rng.uniform(-8, 8, (n, 1)))).astype(np.float32)


@pytest.mark.parametrize("n,n_groups,dim", [(1, 1, 4), (37, 5, 64), (3000, 200, 64), (3000, 700, 16), (5000, 70000, 8),
                                            (2500, 300, 128), (4099, 257, 12), (3000, 2047, 4), (3000, 2048, 4),
                                            (2000, (1 << 22) + 5, 4)])   # 1 / 2 / 3 passes of 11-bit digits
def test_random_groups_bit_exact(n, n_groups, dim):
  rng = np.random.default_rng(n * 31 + dim)
  rows = rows_of(rng, n, dim)
  idx = rng.integers(0, n_groups, size=n).astype(np.int32)
  np.testing.assert_array_equal(reduce_emu(rows, idx, n_groups), O.segment_reduce(rows, idx, n_groups))


@pytest.mark.parametrize("dim,n_huge", [(64, 2), (32, 1), (128, 1), (48, 1), (16, 1), (36, 1)])
def test_huge_groups_are_cut_into_column_slices(dim, n_huge):
  """groups of more than 1024 rows (the Zipf head of a 26 x 65536 batch owns thousands of rows of one key): with rows of
  a multiple of 16 columns every such group becomes dim/16 work items of 16 columns each, pulled from the work queue
  before the ordinary long groups; dims that cannot be sliced (16 itself, 36) keep the whole-row path.  Sums stay in
  position order: bit-identical to the sequential oracle, every output row written."""
  rng = np.random.default_rng(dim * 7 + n_huge)
  n_groups = 60
  parts = [np.full(1100 + 700 * h, 3 + 17 * h) for h in range(n_huge)]        # the huge groups
  parts += [np.full(200, 11), np.full(90, 40)]                                # ordinary long groups
  parts += [rng.integers(0, n_groups, size=900)]                              # short groups and a few empty ones
  idx = np.concatenate(parts).astype(np.int32)
  rng.shuffle(idx)
  rows = rows_of(rng, idx.shape[0], dim)
  got = reduce_emu(rows, idx, n_groups)
  np.testing.assert_array_equal(got, O.segment_reduce(rows, idx, n_groups))


@pytest.mark.parametrize("dim", [64, 7, 36])
def test_zipf_head_takes_the_long_group_path(dim):
  rng = np.random.default_rng(dim)
  n, n_groups = 6000, 400
  idx = np.minimum(rng.zipf(1.3, size=n) - 1, n_groups - 1).astype(np.int32)
  counts = np.bincount(idx, minlength=n_groups)
  assert counts.max() > 500 and (counts == 0).any()      # long groups and empty groups are both present
  rows = rows_of(rng, n, dim)
  np.testing.assert_array_equal(reduce_emu(rows, idx, n_groups), O.segment_reduce(rows, idx, n_groups))


def test_one_group_holds_everything():
  rng = np.random.default_rng(5)
  n, dim = 2100, 20
  rows = rows_of(rng, n, dim)
  idx = np.zeros(n, dtype=np.int32)
  np.testing.assert_array_equal(reduce_emu(rows, idx, 3), O.segment_reduce(rows, idx, 3))


@pytest.mark.parametrize("dim", [132, 1100])
def test_wide_rows(dim):
  rng = np.random.default_rng(dim)
  n, n_groups = 700, 9
  rows = rows_of(rng, n, dim)
  idx = rng.integers(0, n_groups, size=n).astype(np.int32)
  idx[:300] = 4                                            # one long group as well
  np.testing.assert_array_equal(reduce_emu(rows, idx, n_groups), O.segment_reduce(rows, idx, n_groups))


def test_out_of_range_indices_are_dropped():
  rng = np.random.default_rng(9)
  n, n_groups, dim = 1500, 40, 16
  rows = rows_of(rng, n, dim)
  idx = rng.integers(-3, n_groups + 5, size=n).astype(np.int32)
  idx[7] = np.iinfo(np.int32).min
  idx[8] = np.iinfo(np.int32).max
  np.testing.assert_array_equal(reduce_emu(rows, idx, n_groups), O.segment_reduce(rows, idx, n_groups))


def test_empty_input_gives_zero_rows():
  out = reduce_emu(np.zeros((0, 8), np.float32), np.zeros(0, np.int32), 6)
  np.testing.assert_array_equal(out, np.zeros((6, 8), np.float32))


def test_unaligned_rows_take_the_scalar_path():
  rng = np.random.default_rng(11)
  n, n_groups, dim = 900, 50, 16
  buf = np.zeros(n * dim + 1, np.float32)
  rows = buf[1:].reshape(n, dim)                           # 4 B aligned only
  rows[:] = rows_of(rng, n, dim)
  idx = rng.integers(0, n_groups, size=n).astype(np.int32)
  np.testing.assert_array_equal(reduce_emu(rows, idx, n_groups), O.segment_reduce(rows, idx, n_groups))


def test_after_det_unique_like_the_optimizer_path():
  """ids -> det_unique -> per-unique gradient sum: what _resource_apply_sparse_duplicate_indices feeds the optimizer"""
  rng = np.random.default_rng(13)
  n, dim = 4000, 64
  ids = (np.minimum(rng.zipf(1.2, size=n), 900) * 7919 - 5).astype(np.int64)
  u = np.empty(n, np.int64)
  idx = np.empty(n, np.int32)
  cnt = np.zeros(1, np.int64)
  wsb = S().det_unique_workspace_bytes(n)
  ws = np.empty(wsb, np.uint8)
  ck(S().det_unique(P(ids), n, P(u), P(idx), P(cnt), P(ws), wsb, None))
  eu, eidx = O.unique_first_occurrence(ids)
  np.testing.assert_array_equal(idx, eidx)
  g = rows_of(rng, n, dim)
  np.testing.assert_array_equal(reduce_emu(g, idx, int(cnt[0])), O.segment_reduce(g, eidx, eu.shape[0]))


# ---- the bodies of tests/test_segreduce_gpu.py (the Python mirror de.segment_reduce) over the emulated library ----
from tests import test_segreduce_gpu as SG  # noqa: E402
from tests.emu import backend  # noqa: E402


@pytest.fixture
def emu_mirror(monkeypatch):
  with backend.installed():
    monkeypatch.setattr(SG, "DEV", "cpu")
    monkeypatch.setattr(SG, "SCALE", 64)
    yield


@pytest.mark.parametrize("name", ["test_segment_reduce_criteo_shaped_step", "test_segment_reduce_is_deterministic",
                                  "test_segment_reduce_edge_cases"])
def test_gpu_suite_body(emu_mirror, name):
  getattr(SG, name)()


@pytest.mark.parametrize("n,n_groups,dim", [(1000, 37, 16), (3000, 4000, 64), (1500, 50, 260)])
def test_gpu_suite_random_body(emu_mirror, n, n_groups, dim):
  SG.test_segment_reduce_random_bit_exact(n, n_groups, dim)


@pytest.mark.parametrize("mode", ["det", "torch"])
def test_gpu_suite_optimizer_duplicate_ids_body(emu_mirror, monkeypatch, mode):
  SG.test_optimizer_step_with_duplicate_ids(monkeypatch, mode)


@pytest.mark.parametrize("mode", ["det", "torch"])
def test_gpu_suite_lookup_unique_backward_body(emu_mirror, monkeypatch, mode):
  SG.test_embedding_lookup_unique_backward_is_the_gradient_dedupe(monkeypatch, mode)


from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 2500), n_groups=st.integers(1, 70000), dim=st.sampled_from([1, 3, 4, 8, 20, 64, 130]),
       skew=st.sampled_from([0.0, 1.1, 2.0]), seed=st.integers(0, 2**31 - 1))
def test_property_any_shape_matches_the_sequential_sum(n, n_groups, dim, skew, seed):
  rng = np.random.default_rng(seed)
  if skew == 0.0:
    idx = rng.integers(-1, n_groups + 1, size=n)
  else:
    idx = np.minimum(rng.zipf(skew, size=n) - 1, n_groups - 1)
  idx = idx.astype(np.int32)
  rows = rows_of(rng, n, dim)
  np.testing.assert_array_equal(reduce_emu(rows, idx, n_groups), O.segment_reduce(rows, idx, n_groups))


@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
def test_gpu_suite_trainable_sparse_lookup_body(emu_mirror, monkeypatch, combiner):
  SG.test_trainable_sparse_lookup_through_the_fused_segment_sum(monkeypatch, combiner)
