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
def _interleaved_repair_round(m, rng):
  """threads = displaced keys; each runs  check -> claim (CAS on the first free slot of its chain) -> copy -> erase
  with arbitrary interleaving (the kernel: one lane per slot, claims via warp_find_or_claim)."""
  threads = []
  for s in range(m.nb * BUCKET):
    k = m.keys[s]
    if k not in (EMPTY, TOMB) and m.home(k) != s // BUCKET:
      threads.append({"s": s, "k": k, "pc": 0, "dst": -1})
  moves = 0
  live = list(range(len(threads)))
  while live:
    i = rng.choice(live)
    t = threads[i]
    if t["pc"] == 0:     # reachability check (a snapshot that may be stale by the time of the claim)
      t["pc"] = 1 if not m.reachable(t["s"]) else 9
    elif t["pc"] == 1:   # find-or-claim along the chain with fresh loads
      found, first_free = -1, -1
      for _, chain_slots in m._chain(t["k"]):
        for q in chain_slots:
          if m.keys[q] == t["k"]:
            found = q
          if first_free < 0 and m.keys[q] in (EMPTY, TOMB):
            first_free = q
      if found >= 0:
        t["pc"] = 9      # reachable again (somebody filled the gap): nothing to do
      else:
        assert first_free >= 0
        if m.keys[first_free] == TOMB:
          m.used -= 1
        m.keys[first_free] = t["k"]   # CAS succeeded (a failed CAS restarts this step)
        t["dst"] = first_free
        t["pc"] = 2
    elif t["pc"] == 2:   # row + score copy
      m.vals[t["dst"]] = m.vals[t["s"]]
      m.scores[t["dst"]] = m.scores[t["s"]]
      t["pc"] = 3
    elif t["pc"] == 3:   # erase the old slot
      m.keys[t["s"]] = EMPTY
      m.scores[t["s"]] = 0
      m.vals.pop(t["s"], None)
      moves += 1
      t["pc"] = 9
    if t["pc"] == 9:
      live.remove(i)
  return moves


@pytest.mark.parametrize("seed", range(12))
def test_repair_under_random_interleavings(seed):
  rng = random.Random(seed)
  m = EvictModel(48, CUSTOMIZED, max_lf=0.9)
  keys = rng.sample(range(1 << 40), m.limit)
  scores = [rng.randrange(0, 1000) for _ in keys]
  m.insert_scored(keys, [k ^ 5 for k in keys], scores)
  m.check_invariants()
  content = m.live()
  k = len(keys) // 3
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
  assert m.used == m.size       # eviction leaves no tombstones behind
