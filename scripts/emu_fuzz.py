"""CPU, no GPU: randomised differential runs of the kernels written after round 1's GPU budget was spent, on the SIMT
emulator against the oracle -- many more cases than the test suite's property tests.
  python scripts/emu_fuzz.py [cases]      (default 400: det_segment_reduce, 250: staged segment-sum)"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.test_segreduce_emu import reduce_emu, rows_of  # noqa: E402
from tests.test_segsum_staged_emu import _both, _tables  # noqa: E402


def fuzz_segment_reduce(cases):
  rng0 = np.random.default_rng(12345)
  bad = 0
  for it in range(cases):
    seed = int(rng0.integers(0, 2**31))
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 6000))
    n_groups = int(rng.choice([1, 2, 7, 255, 256, 257, 1000, 65535, 65536, 70000, 300000]))
    dim = int(rng.choice([1, 2, 3, 4, 8, 12, 16, 20, 36, 64, 100, 128, 132, 260]))
    skew = float(rng.choice([0.0, 1.05, 1.3, 2.5]))
    idx = rng.integers(-2, n_groups + 2, size=n) if skew == 0.0 else np.minimum(rng.zipf(skew, size=n) - 1, n_groups - 1)
    idx = idx.astype(np.int32)
    rows = rows_of(rng, n, dim)
    if not np.array_equal(reduce_emu(rows, idx, n_groups), O.segment_reduce(rows, idx, n_groups)):
      bad += 1
      print("MISMATCH det_segment_reduce", seed, n, n_groups, dim, skew)
  print("det_segment_reduce:", cases, "cases, mismatches:", bad)
  return bad


def fuzz_staged_segment_sum(cases):
  rng0 = np.random.default_rng(777)
  bad = 0
  for it in range(cases):
    seed = int(rng0.integers(0, 2**31))
    rng = np.random.default_rng(seed)
    dim = int(rng.choice([4, 8, 12, 16, 32, 48, 64, 96, 128]))
    batch = int(rng.integers(1, 1500))
    maxlen = int(rng.choice([1, 2, 5, 20, 70, 300]))
    p_empty = float(rng.choice([0.0, 0.2, 0.7, 0.97]))
    combiner = str(rng.choice(["sum", "mean", "sqrtn"]))
    use_w = bool(rng.integers(0, 2))
    t, ot = _tables(rng, dim)
    lens = rng.integers(1, maxlen + 1, size=batch) * (rng.random(batch) >= p_empty)
    seg = np.repeat(np.arange(batch), lens).astype(np.int32)
    ids = rng.integers(0, 1100, size=seg.shape[0]).astype(np.int64)   # some ids are not in the table: default row
    w = rng.uniform(0.25, 2.0, size=seg.shape[0]).astype(np.float32) if use_w else None
    mp = pytest.MonkeyPatch()
    try:
      _both(mp, t, ot, ids, seg, w, batch, combiner)
    except AssertionError as e:
      bad += 1
      print("MISMATCH staged segment-sum", seed, dim, batch, maxlen, p_empty, combiner, use_w, str(e)[:200])
    finally:
      mp.undo()
      t.close()
  print("segment_sum_staged_kernel:", cases, "cases, mismatches:", bad)
  return bad


if __name__ == "__main__":
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
  sys.exit(1 if fuzz_segment_reduce(n) + fuzz_staged_segment_sum(max(1, n * 5 // 8)) else 0)
