"""CPU tests of the score-based capacity management (the model of csrc/evict.cu's host algorithm, tests/evict_model.py):
(1) the reference's own eviction tests (kernel_tests/hkv_hashtable_evict_test.py:110-577) replayed on the model --
the same bodies run against the CUDA path in tests/test_evict_gpu.py; (2) invariants under arbitrary op streams;
(3) a random-interleaving simulation of the parallel repair pass."""
import random

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as hst

from tests.evict_model import CUSTOMIZED, EPOCHLFU, EPOCHLRU, LFU, LRU, EvictModel
from tests.layout_model import BUCKET, EMPTY, TOMB

STRATEGIES = [LRU, LFU, EPOCHLRU, EPOCHLFU, CUSTOMIZED]
DIM = 8


class HkvModelTable(object):
  """gpu::TableWrapper::upsert (lookup_table_op_hkv.h:519-535: epoch stepping) + HkvHashTable._gen_scores
  (python/ops/hkv_hashtable_ops.py:209-216) over the model -- what de.HkvHashTable does over the C ABI."""

  def __init__(self, strategy, capacity=1024, step_per_epoch=0, gen_scores_fn=None):
    self.m = EvictModel(capacity // BUCKET, strategy)
    self.strategy = strategy
    self.step_per_epoch = step_per_epoch
    self.gen_scores_fn = gen_scores_fn
    self.curr_epoch, self.curr_step = 0, 1

  def upsert(self, keys, values):
    keys = np.asarray(keys, dtype=np.int64)
    if self.strategy == CUSTOMIZED:
      scores = self.gen_scores_fn(keys)
    elif self.strategy in (LFU, EPOCHLFU):
      scores = np.ones(len(keys), dtype=np.int64)
    else:
      scores = None
    self.m.epoch = self.curr_epoch
    self.m.insert_scored(list(keys), list(values), scores)
    if self.strategy in (EPOCHLRU, EPOCHLFU):
      self.curr_step += 1
      if self.curr_step > self.step_per_epoch:
        self.curr_epoch += 1
        self.curr_step = 1
    self.m.check_invariants()

  def lookup(self, keys):
    out = []
    for k in keys:
      s = self.m.find(int(k))
      out.append(self.m.vals[s] if s >= 0 else [0] * DIM)
    return np.array(out)

  def export_keys_and_scores(self):
    ks, sc = self.m.export_keys_and_scores()
    return ks, sc.astype(np.int64)


def gen_scores_fn(keys):
  return np.asarray(keys, dtype=np.int64) + 1


def rows(vals):
  return [[v] * DIM for v in vals]


# ---- (1) the reference's tests ---------------------------------------------------------------------------
@pytest.mark.parametrize("strategy", STRATEGIES)
def test_evict_strategy_basic_and_export_scores(strategy):
  """hkv_hashtable_evict_test.py:110-230 (test_evict_strategy, test_export_keys_and_scores)"""
  t = HkvModelTable(strategy, gen_scores_fn=gen_scores_fn, step_per_epoch=4)
  keys = np.array([0, 1, 2, 3])
  t.upsert(keys, rows([0, 1, 2, 3]))
  assert (t.lookup(keys) == np.array(rows([0, 1, 2, 3]))).all()
  ek, es = t.export_keys_and_scores()
  assert (np.sort(ek) == keys).all()
  if strategy == CUSTOMIZED:
    assert (np.sort(es) == gen_scores_fn(keys)).all()
  elif strategy in (EPOCHLFU, LFU):
    assert (es == 1).all()


def test_evict_strategy_lfu():
  """:232-309"""
  t = HkvModelTable(LFU)
  keys = np.array([0, 1, 2, 3])
  t.upsert(keys, rows([0, 1, 2, 3]))
  assert (t.export_keys_and_scores()[1] == 1).all()
  t.upsert(keys, rows([0, 1, 2, 3]))
  assert (t.export_keys_and_scores()[1] == 2).all()
  t.upsert(np.array([0, 1, 4, 5]), rows([0, 1, 2, 3]))
  assert (np.sort(t.export_keys_and_scores()[1]) == np.array([1, 1, 2, 2, 3, 3])).all()
  keys = np.arange(4, 1034)
  t.upsert(keys, rows([10] * len(keys)))
  ek, es = t.export_keys_and_scores()
  assert len(ek) < 1024
  assert (np.sort(ek)[:6] == np.arange(0, 6)).all()
  assert (np.sort(es)[-6:] == np.array([2, 2, 2, 2, 3, 3])).all()


def test_evict_strategy_epoch_lfu():
  """:311-405"""
  t = HkvModelTable(EPOCHLFU, step_per_epoch=4)
  for base in [1, 1 + (1 << 32), 1 + (2 << 32)]:
    keys = np.array([0, 1, 2, 3])
    t.upsert(keys, rows([0, 1, 2, 3]))
    es = t.export_keys_and_scores()[1]
    assert (np.sort(es)[-4:] >= base).all()
    t.upsert(keys, rows([0, 1, 2, 3]))
    es = t.export_keys_and_scores()[1]
    assert (np.sort(es)[-4:] >= base + 1).all()
    t.upsert(np.array([0, 1, 4, 5]), rows([0, 1, 2, 3]))
    es = t.export_keys_and_scores()[1]
    assert (np.sort(es)[-6:] >= np.array([base, base, base + 1, base + 1, base + 2, base + 2])).all()
    keys = np.arange(4, 1024)
    t.upsert(keys, rows([10] * len(keys)))
    ek, es = t.export_keys_and_scores()
    assert len(ek) < 1024
    assert (np.sort(ek)[:6] == np.arange(0, 6)).all()
    assert (np.sort(es)[-6:] >= np.array([base + 1] * 4 + [base + 2] * 2)).all()


def test_evict_strategy_lru():
  """:407-476"""
  t = HkvModelTable(LRU)
  keys = np.array([0, 1, 2, 3])
  t.upsert(keys, rows([0, 1, 2, 3]))
  assert np.isin(keys, t.export_keys_and_scores()[0]).all()
  t.upsert(np.array([2, 3, 6, 7]), rows([0, 1, 2, 3]))
  ek, es = t.export_keys_and_scores()
  l1 = [int(s) for k, s in zip(ek, es) if k in (0, 1)]
  l2 = [int(s) for k, s in zip(ek, es) if k in (2, 3)]
  assert l1 < l2
  keys = np.arange(4, 1044)
  t.upsert(keys, rows([10] * len(keys)))
  keys = np.arange(1024, 1400)
  t.upsert(keys, rows([10] * len(keys)))
  ek, _ = t.export_keys_and_scores()
  assert len(ek) <= 1024
  assert not np.isin(np.arange(0, 4), ek).any()


def test_evict_strategy_epoch_lru():
  """:478-519"""
  t = HkvModelTable(EPOCHLRU, step_per_epoch=1)
  for epoch in range(2):
    keys = np.arange(0, 1024)
    t.upsert(keys, rows([10] * len(keys)))
    _, es = t.export_keys_and_scores()
    assert (es >= (epoch << 32)).all()
    assert (es < (epoch << 32) + 0xffffffff).all()


def test_evict_strategy_custom():
  """:521-573: keys offered with a score below every resident score are refused"""
  calls = [0]

  def gen_custom(keys):
    calls[0] += 1
    return np.full(len(keys), 10000 if calls[0] == 1 else 1, dtype=np.int64)

  t = HkvModelTable(CUSTOMIZED, gen_scores_fn=gen_custom)
  keys = np.arange(2048, 4096)
  t.upsert(keys, rows([10] * len(keys)))
  keys = np.arange(0, 1024)
  t.upsert(keys, rows([10] * len(keys)))
  ek, es = t.export_keys_and_scores()
  assert len(ek) > 0
  assert (es == 10000).all()
  assert (ek >= 1024).all()
  assert t.m.n_refused == 1024


# ---- (2) invariants -----------------------------------------------------------------------------------------
def test_evict_lowest_takes_exactly_the_k_lowest():
  rng = random.Random(5)
  m = EvictModel(64, CUSTOMIZED)
  keys = rng.sample(range(1 << 40), 400)
  scores = [rng.randrange(0, 50) for _ in keys]   # many ties
  m.insert_scored(keys, [0] * len(keys), scores)
  by_key = dict(zip(keys, scores))
  for k in (1, 7, 100, 250):
    before = sorted(by_key.values())
    ev = m.evict_lowest(k)
    m.check_invariants()
    assert len(ev) == k
    assert sorted(by_key[x] for x in ev) == before[:k]
    for x in ev:
      del by_key[x]
    assert set(m.live().keys()) == set(by_key.keys())


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(hst.sampled_from(STRATEGIES), hst.integers(4, 24), hst.lists(
    hst.tuples(hst.sampled_from(["ins", "ins", "ins", "rem"]), hst.lists(hst.integers(0, 400), min_size=1, max_size=60,
                                                                       unique=True)), min_size=1, max_size=30))
def test_invariants_under_random_streams(strategy, nb, ops):
  m = EvictModel(nb, strategy)
  cap = nb * BUCKET
  for op, ks in ops:
    if op == "ins":
      refused = m.insert_scored(ks, [k * 3 for k in ks], [k % 7 for k in ks] if strategy == CUSTOMIZED else None)
      for k in ks:
        if k in refused:
          assert m.find(k) < 0
      # the keys of the LAST chunk that were admitted are resident with the value just written
      chunk = max(1, m.limit // 4)
      for k in ks[(len(ks) - 1) // chunk * chunk:]:
        if k not in refused:
          s = m.find(k)
          assert s >= 0 and m.vals[s] == k * 3
    else:
      for k in ks:
        m.remove(k)
    m.check_invariants()
    assert m.size <= cap and m.used <= cap


# ---- (3) the repair pass under random interleavings ------------------------------------------------------------
def _probe_steps(m, key):
  """generator: walks the probe chain of `key` one bucket per step (other threads run in between), yields None
  while walking and finally ('done', first match or -1, first free slot or -1)"""
  b = m.home(key)
  found, first_free = -1, -1
  for _ in range(m.nb):
    lo = b * BUCKET
    ks = [m.keys[q] for q in range(lo, lo + BUCKET)]       # one 64 B bucket load
    for j, k in enumerate(ks):
      if found < 0 and k == key:
        found = lo + j
      if first_free < 0 and k in (EMPTY, TOMB):
        first_free = lo + j
    if found >= 0 or EMPTY in ks:
      break
    b = (b + 1) % m.nb
    yield None
  yield ("done", found, first_free)


def _interleaved_repair_round(m, rng):
  """the two kernels of a repair round with arbitrary interleaving of their threads (one thread per displaced key;
  the kernels themselves are separated by a launch boundary).
  move : check -> probe bucket by bucket -> CAS claim -> copy            (EMPTY slots only disappear)
  sweep: probe bucket by bucket -> erase when an earlier match exists     (EMPTY slots only appear)"""
  work = 0
  threads = [{"s": s, "k": m.keys[s], "pc": 0} for s in range(m.nb * BUCKET)
             if m.keys[s] not in (EMPTY, TOMB) and m.home(m.keys[s]) != s // BUCKET]
  live = list(range(len(threads)))
  while live:
    i = rng.choice(live)
    t = threads[i]
    if t["pc"] == 0:
      if m.reachable(t["s"]):
        t["pc"] = 9
      else:
        t["gen"] = _probe_steps(m, t["k"])
        t["pc"] = 1
    elif t["pc"] == 1:
      r = next(t["gen"])
      if r is not None:
        _, found, first_free = r
        if found >= 0:
          t["pc"] = 9                      # reachable again, or a duplicate's mover got there first
        else:
          assert first_free >= 0
          if m.keys[first_free] in (EMPTY, TOMB):   # the CAS
            if m.keys[first_free] == EMPTY:
              m.used += 1
            m.keys[first_free] = t["k"]
            t["dst"] = first_free
            t["pc"] = 2
          else:
            t["gen"] = _probe_steps(m, t["k"])      # CAS lost: rescan the chain
    elif t["pc"] == 2:
      m.vals[t["dst"]] = m.vals[t["s"]]
      m.scores[t["dst"]] = m.scores[t["s"]]
      work += 1
      t["pc"] = 9
    if t["pc"] == 9:
      live.remove(i)
  # ---- launch boundary ----
  threads = [{"s": s, "k": m.keys[s], "gen": None} for s in range(m.nb * BUCKET)
             if m.keys[s] not in (EMPTY, TOMB) and m.home(m.keys[s]) != s // BUCKET]
  live = list(range(len(threads)))
  while live:
    i = rng.choice(live)
    t = threads[i]
    if t["gen"] is None:
      t["gen"] = _probe_steps(m, t["k"])
    r = next(t["gen"])
    if r is None:
      continue
    _, found, _ff = r
    if found >= 0 and found != t["s"]:
      m.keys[t["s"]] = EMPTY
      m.scores[t["s"]] = 0
      m.vals.pop(t["s"], None)
      m.used -= 1
      work += 1
    live.remove(i)
  return work


@pytest.mark.parametrize("seed", range(24))
def test_repair_under_random_interleavings(seed):
  rng = random.Random(seed)
  m = EvictModel(rng.choice([8, 16, 48]), CUSTOMIZED, max_lf=rng.choice([0.8, 0.9, 0.95]))
  keys = rng.sample(range(1 << 40), m.limit)
  scores = [rng.randrange(0, 1000) for _ in keys]
  m.insert_scored(keys, [k ^ 5 for k in keys], scores)
  if seed % 2:                       # user removes leave tombstones in full buckets: movers recycle them
    for k in rng.sample(keys, len(keys) // 10):
      m.remove(k)
  m.check_invariants()
  content = m.live()
  k = max(1, len(content) // rng.choice([2, 3, 10]))
  tau, quota = m.select_threshold(k)
  ev = m.evict_apply(tau, quota)
  for x in ev:
    del content[x]
  rounds = 0
  while _interleaved_repair_round(m, rng):
    rounds += 1
    assert rounds < 64
  m.check_invariants()          # every key reachable again, stored once, free slots carry score 0
  assert m.live() == content    # nothing lost, rows travelled with their keys
  if not seed % 2:
    assert m.used == m.size     # eviction leaves no tombstones behind
