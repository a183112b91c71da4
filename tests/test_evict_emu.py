"""The REAL device code of csrc/evict_kernels.cuh + csrc/common.cuh (probe / claim primitives, scored insert, classify,
min/max, radix select, evict-apply, the two repair kernels, score carry) executed on the CPU under the SIMT emulator
of tests/emu/ (every CUDA thread an OS thread, warp collectives as barriers, atomics as real atomics), and checked
against the independent Python restatement of the layout and of the eviction algorithm (tests/layout_model.py,
tests/evict_model.py).  No GPU needed; the emulator is test infrastructure and never part of the product."""
import ctypes

import numpy as np
import pytest

from tests.emu import build_emu
from tests.evict_model import CUSTOMIZED, EPOCHLFU, EPOCHLRU, LFU, LRU, EvictModel
from tests.layout_model import BUCKET, EMPTY, TOMB

_vp, _u64, _sz, _i = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_int
_LIB = None


def lib():
  global _LIB
  if _LIB is None:
    l = ctypes.CDLL(build_emu.build())
    l.emu_create.restype = _vp
    l.emu_create.argtypes = [_u64, _i, _i, _i]
    l.emu_destroy.argtypes = [_vp]
    l.emu_set_epoch.argtypes = [_vp, _u64]
    l.emu_state.argtypes = [_vp, _vp, _vp, _vp]
    l.emu_dump.argtypes = [_vp, _vp, _vp, _vp]
    l.emu_slot_plane.argtypes = [_vp, _i, _vp]
    l.emu_insert_scored.argtypes = [_vp, _vp, _vp, _vp, _vp, _sz, _i]
    l.emu_touch.argtypes = [_vp, _vp, _vp, _sz, _i]
    l.emu_scores_of.argtypes = [_vp, _vp, _sz, _vp, _i, _i]
    l.emu_find.argtypes = [_vp, _vp, _sz, _vp, _i]
    l.emu_classify.argtypes = [_vp, _vp, _vp, _sz, _i, _vp, _vp, _vp, _i]
    l.emu_evict_lowest.restype = _u64
    l.emu_evict_lowest.argtypes = [_vp, _u64, _i, _vp, _vp, _vp]
    l.emu_carry_scores.argtypes = [_vp, _vp, _i]
    l.emu_remove_slots.argtypes = [_vp, _vp, _sz]
    _LIB = l
  return _LIB


def P(a):
  return None if a is None else a.ctypes.data_as(_vp)


class EmuTable(object):

  def __init__(self, nb, dim, strategy, n_slot_planes=0, grid=2):
    self.nb, self.dim, self.strategy, self.grid, self.np = nb, dim, strategy, grid, n_slot_planes
    self.cap = nb * BUCKET
    self.h = lib().emu_create(nb, dim, n_slot_planes, strategy)

  def __del__(self):
    if getattr(self, "h", None):
      lib().emu_destroy(self.h)
      self.h = None

  def set_epoch(self, e):
    lib().emu_set_epoch(self.h, e)

  def insert(self, keys, values, scores=None, mask=None):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    values = np.ascontiguousarray(values, dtype=np.float32).reshape(len(keys), self.dim)
    scores = None if scores is None else np.ascontiguousarray(scores, dtype=np.uint64)
    mask = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    lib().emu_insert_scored(self.h, P(keys), P(values), P(scores), P(mask), len(keys), self.grid)

  def touch(self, keys, scores=None):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    scores = None if scores is None else np.ascontiguousarray(scores, dtype=np.uint64)
    lib().emu_touch(self.h, P(keys), P(scores), len(keys), self.grid)

  def find(self, keys):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    out = np.empty(len(keys), dtype=np.int64)
    lib().emu_find(self.h, P(keys), len(keys), P(out), self.grid)
    return out

  def scores_of(self, keys):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    out = np.empty(len(keys), dtype=np.uint64)
    lib().emu_scores_of(self.h, P(keys), len(keys), P(out), 0, self.grid)
    return out

  def remove(self, keys):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    lib().emu_scores_of(self.h, P(keys), len(keys), None, 1, self.grid)   # evict_before_remove
    slots = self.find(keys)
    lib().emu_remove_slots(self.h, P(slots), len(slots))

  def classify(self, keys, scores, admission):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    scores = None if scores is None else np.ascontiguousarray(scores, dtype=np.uint64)
    mask = np.zeros(len(keys), dtype=np.uint8)
    n_new, n_adm = _u64(0), _u64(0)
    lib().emu_classify(self.h, P(keys), P(scores), len(keys), int(admission), P(mask), ctypes.byref(n_new),
                       ctypes.byref(n_adm), self.grid)
    return mask, n_new.value, n_adm.value

  def evict_lowest(self, k):
    rounds, tau, quota = _i(0), _u64(0), _u64(0)
    n = lib().emu_evict_lowest(self.h, k, self.grid, ctypes.byref(rounds), ctypes.byref(tau), ctypes.byref(quota))
    return n, rounds.value, tau.value, quota.value

  def state(self):
    size, used, err = _u64(0), _u64(0), ctypes.c_uint(0)
    lib().emu_state(self.h, ctypes.byref(size), ctypes.byref(used), ctypes.byref(err))
    return size.value, used.value, err.value

  def dump(self):
    keys = np.empty(self.cap, dtype=np.int64)
    scores = np.empty(self.cap + 2, dtype=np.uint64)
    rows = np.empty((self.cap + 2, self.dim), dtype=np.float32)
    lib().emu_dump(self.h, P(keys), P(scores), P(rows))
    return keys, scores, rows

  def as_model(self):
    """the raw planes loaded into the Python model: its invariant checker (independent restatement of hashing,
    probe chains, counters, free-slots-carry-score-0) then judges what the device code produced"""
    keys, scores, rows = self.dump()
    m = EvictModel(self.nb, self.strategy)
    m.keys = np.array([int(k) for k in keys], dtype=object)
    m.scores = [int(s) for s in scores]
    m.vals = {s: rows[s].copy() for s in range(self.cap) if int(keys[s]) not in (EMPTY, TOMB)}
    m.size, m.used, err = self.state()
    assert err == 0
    return m

  def content(self):
    keys, scores, rows = self.dump()
    return {int(keys[s]): (int(scores[s]), rows[s].copy()) for s in range(self.cap) if int(keys[s]) not in (EMPTY, TOMB)}


def rows_for(keys, dim):
  keys = np.asarray(keys, dtype=np.int64)
  return ((keys[:, None] % 1000).astype(np.float32) + np.arange(dim, dtype=np.float32)[None, :] / 64.0)


@pytest.mark.parametrize("dim", [4, 6])        # 16 B vector rows / 4 B vector rows
def test_scored_insert_matches_the_layout_model(dim):
  rng = np.random.default_rng(1)
  t = EmuTable(96, dim, CUSTOMIZED, n_slot_planes=1)
  keys = rng.choice(1 << 40, size=600, replace=False).astype(np.int64)
  scores = rng.integers(0, 1 << 40, size=len(keys)).astype(np.uint64)
  t.insert(keys, rows_for(keys, dim), scores)
  m = t.as_model()
  m.check_invariants()
  assert m.size == len(keys)
  c = t.content()
  assert set(c) == set(keys.tolist())
  for k, s, r in zip(keys.tolist(), scores.tolist(), rows_for(keys, dim)):
    assert c[k][0] == s and np.array_equal(c[k][1], r)
  # slots found by the device probe are where the model's probe finds them
  slots = t.find(np.concatenate([keys[:100], np.array([-5, 12345678901], dtype=np.int64)]))
  assert [m.find(int(k)) for k in keys[:100]] == slots[:100].tolist() and slots[100:].tolist() == [-1, -1]
  # re-assign half of them: rows and scores are overwritten, nothing new appears
  t.insert(keys[:300], rows_for(keys[:300], dim) + 1, scores[:300] + np.uint64(7))
  c = t.content()
  assert len(c) == len(keys) and all(c[int(k)][0] == int(s) + 7 for k, s in zip(keys[:300], scores[:300]))
  t.as_model().check_invariants()


def test_duplicates_in_one_launch_are_stored_once():
  rng = np.random.default_rng(2)
  t = EmuTable(32, 4, LFU)
  base = rng.choice(1 << 30, size=60, replace=False).astype(np.int64)
  keys = np.concatenate([base, base, base[:30]])
  rng.shuffle(keys)
  t.insert(keys, rows_for(keys, 4))
  m = t.as_model()
  m.check_invariants()
  assert m.size == len(base)


def test_score_rules():
  keys = np.arange(100, 140, dtype=np.int64)
  t = EmuTable(16, 4, LFU)
  t.insert(keys, rows_for(keys, 4))
  t.insert(keys[:10], rows_for(keys[:10], 4), np.full(10, 5, dtype=np.uint64))
  t.touch(keys[5:15])
  sc = t.scores_of(keys)
  assert sc[:5].tolist() == [6] * 5 and sc[5:10].tolist() == [7] * 5 and sc[10:15].tolist() == [2] * 5
  assert (sc[15:] == 1).all() and t.scores_of(np.array([7], dtype=np.int64))[0] == 0
  t = EmuTable(16, 4, EPOCHLFU)
  t.insert(keys, rows_for(keys, 4))
  t.set_epoch(3)
  t.insert(keys[:4], rows_for(keys[:4], 4))
  sc = t.scores_of(keys)
  assert sc[:4].tolist() == [(3 << 32) | 2] * 4 and (sc[4:] == 1).all()
  t = EmuTable(16, 4, LRU)
  t.insert(keys[:20], rows_for(keys[:20], 4))
  t.insert(keys[20:], rows_for(keys[20:], 4))
  sc = t.scores_of(keys)
  assert sc[:20].max() < sc[20:].min()
  t = EmuTable(16, 4, EPOCHLRU)
  t.set_epoch(2)
  t.insert(keys, rows_for(keys, 4))
  sc = t.scores_of(keys)
  assert ((sc >> np.uint64(32)) == 2).all()


def test_classify_counts_and_admission_mask():
  rng = np.random.default_rng(3)
  t = EmuTable(32, 4, CUSTOMIZED)
  resident = rng.choice(1 << 30, size=150, replace=False).astype(np.int64)
  t.insert(resident, rows_for(resident, 4), rng.integers(100, 200, size=150).astype(np.uint64))
  smin = int(t.scores_of(resident).min())
  new = (np.arange(80, dtype=np.int64) + (1 << 31))
  batch = np.concatenate([resident[:40], new])
  scores = np.concatenate([np.zeros(40, dtype=np.uint64), rng.integers(50, 250, size=80).astype(np.uint64)])
  mask, n_new, n_adm = t.classify(batch, scores, admission=True)
  assert n_new == 80 and n_adm == int((scores[40:] >= smin).sum())
  assert mask[:40].all() and (mask[40:] == (scores[40:] >= smin)).all()
  mask, n_new, n_adm = t.classify(batch, scores, admission=False)
  assert n_new == 80 and n_adm == 80 and mask.all()
  # the may-claim mask keeps refused keys out, resident keys are still assigned
  t.insert(batch, rows_for(batch, 4), scores, mask=t.classify(batch, scores, admission=True)[0])
  c = t.content()
  assert all((int(k) in c) == bool(s >= smin) for k, s in zip(new, scores[40:]))
  t.as_model().check_invariants()


@pytest.mark.parametrize("seed,tombstones", [(0, False), (1, False), (2, True), (3, True)])
def test_eviction_event_matches_the_model(seed, tombstones):
  """select (threshold, tie quota) == the model's, the k lowest go, survivors keep rows / scores / optimizer slots and
  stay reachable after the repair rounds, no tombstones are created"""
  rng = np.random.default_rng(seed)
  nb, dim = 64, 4
  t = EmuTable(nb, dim, CUSTOMIZED, n_slot_planes=1)
  n = int(nb * BUCKET * 0.9)
  keys = rng.choice(1 << 40, size=n, replace=False).astype(np.int64)
  scores = rng.integers(0, 300, size=n).astype(np.uint64)       # many ties
  if seed == 1:
    scores = (scores << np.uint64(40)) | rng.integers(0, 1 << 20, size=n).astype(np.uint64)   # wide scores: 6 passes
  t.insert(keys, rows_for(keys, dim), scores)
  if tombstones:
    t.remove(keys[::9])
  m = t.as_model()
  m.check_invariants()
  before = t.content()
  for k_ev in (1, len(before) // 3):
    tau_m, quota_m = m.select_threshold(min(k_ev, len(before)))
    n_ev, rounds, tau, quota = t.evict_lowest(k_ev)
    assert (tau, quota) == (tau_m, quota_m)
    assert n_ev == k_ev
    after = t.content()
    gone = set(before) - set(after)
    assert len(gone) == k_ev and set(after) <= set(before)
    assert all(before[k][0] <= tau for k in gone) and all(before[k][0] >= tau for k in after)
    assert sum(1 for k in gone if before[k][0] == tau) == quota
    for k, (s, r) in after.items():
      assert s == before[k][0] and np.array_equal(r, before[k][1])
    m = t.as_model()
    m.check_invariants()
    if not tombstones:
      assert m.used == m.size
    before = after


def test_steady_state_on_device_code_follows_the_model_policy():
  """the host policy of evict_room / evict_insert (tests/evict_model.py: chunks, soft / hard limits, admission, slab)
  driven step by step over the device kernels: LFU churn at the limit"""
  rng = np.random.default_rng(7)
  nb, dim = 48, 4
  t = EmuTable(nb, dim, LFU, grid=1)
  cap = nb * BUCKET
  limit, hard = int(cap * 0.875), int(cap * 0.95)
  slab = max(1, limit // 32)
  hot = np.arange(10, dtype=np.int64)
  nxt = 1000
  events = 0
  for step in range(14):
    new = np.arange(nxt, nxt + 60, dtype=np.int64)
    nxt += 60
    batch = np.concatenate([hot, new])
    size, used, _ = t.state()
    mask = None
    if used + len(batch) > limit:
      mask, n_new, n_adm = t.classify(batch, None, admission=True)
      if used + n_adm > limit:
        need = used + n_adm - limit
        n_ev, _, _, _ = t.evict_lowest(min(size, need + slab))
        events += 1
        assert n_ev == min(size, need + slab)
    t.insert(batch, rows_for(batch, dim), None, mask)
    size, used, err = t.state()
    assert used <= hard and err == 0
  assert events >= 3
  m = t.as_model()
  m.check_invariants()
  c = t.content()
  assert all(int(k) in c and c[int(k)][0] == 14 for k in hot)      # the hot keys were never evicted
  assert m.used == m.size


def test_scores_follow_their_keys_into_a_bigger_table():
  rng = np.random.default_rng(9)
  old = EmuTable(16, 4, CUSTOMIZED)
  keys = rng.choice(1 << 30, size=100, replace=False).astype(np.int64)
  scores = rng.integers(1, 1 << 50, size=100).astype(np.uint64)
  old.insert(keys, rows_for(keys, 4), scores)
  new = EmuTable(40, 4, CUSTOMIZED)
  new.insert(keys, rows_for(keys, 4), np.zeros(100, dtype=np.uint64))     # what rehash_kernel leaves: keys + rows
  lib().emu_carry_scores(old.h, new.h, 2)
  assert new.scores_of(keys).tolist() == scores.tolist()
  new.as_model().check_invariants()
