"""Invariants of the GPU table's layout logic, checked on the sequential model (no GPU): a key is stored at most
once and stays reachable along its probe chain under arbitrary insert / overwrite / remove streams, including
full buckets, tombstone recycling, the erase-to-EMPTY shortcut and the two sentinel-valued keys."""
import numpy as np
import pytest

from tests.layout_model import EMPTY, TOMB, LayoutModel

try:
  from hypothesis import given, settings, strategies as st
  HAVE = True
except Exception:  # pragma: no cover
  HAVE = False


def test_dense_fill_and_churn_small_table():
  rng = np.random.default_rng(0)
  m, ref = LayoutModel(nb=4), {}            # 32 slots: buckets overflow into their neighbours all the time
  universe = list(range(-30, 60)) + [EMPTY, TOMB]
  for step in range(4000):
    k = universe[int(rng.integers(len(universe)))]
    if rng.random() < 0.55 and (len(ref) < 24 or k in ref):   # load factor <= 0.75 like the engine enforces
      m.insert(k, step)
      ref[k] = step
    else:
      m.remove(k)
      ref.pop(k, None)
    if step % 97 == 0:
      m.check_invariants()
  m.check_invariants()
  assert m.live() == ref
  for k in universe:
    assert (m.find(k) >= 0) == (k in ref)


if HAVE:
  _ops = st.lists(st.tuples(st.sampled_from(["ins", "del"]), st.integers(min_value=-12, max_value=12)), max_size=200)

  @settings(max_examples=150, deadline=None)
  @given(st.sampled_from([1, 2, 3, 5]), _ops)
  def test_layout_invariants_property(nb, ops):
    m, ref = LayoutModel(nb), {}
    cap = int(nb * 8 * 0.75)
    for i, (op, k) in enumerate(ops):
      if op == "ins":
        if k not in ref and len(ref) >= cap:
          continue                        # the engine grows before this point; the model keeps its size
        m.insert(k, i)
        ref[k] = i
      else:
        m.remove(k)
        ref.pop(k, None)
    m.check_invariants()
    assert m.live() == ref


def test_erase_to_empty_shortcut_never_cuts_a_chain():
  """The shortcut writes EMPTY only into a bucket that already has an EMPTY slot; such a bucket ends every chain
  through it anyway, so no key placed further along can become unreachable."""
  m = LayoutModel(nb=3)
  # fill bucket of key 0 completely so that later keys of that home bucket spill into the next bucket
  home0 = m.home(0)
  same = [k for k in range(0, 4000) if m.home(k) == home0][:11]
  for k in same:
    m.insert(k, k)
  spilled = [k for k in same if m.find(k) // 8 != home0]
  assert spilled, "test needs keys that overflowed into the next bucket"
  m.remove(same[0])                       # home bucket was full -> TOMBSTONE, not EMPTY
  assert m.keys[home0 * 8:(home0 + 1) * 8].tolist().count(TOMB) == 1
  for k in spilled:
    assert m.find(k) >= 0
  m.insert(99991 if m.home(99991) == home0 else same[0], -1)   # recycles the tombstone
  m.check_invariants()


# ---- the concurrent find-or-claim protocol under arbitrary interleavings -----------------------------------
class _Agent(object):
  """One 4-lane subgroup of warp_find_or_claim_t as a step machine: LOAD a bucket view, PROCESS it, CAS."""

  def __init__(self, model, key, prefetched):
    self.m, self.key = model, key
    self.b0 = model.home(key)
    self.b, self.first_free, self.ff_tomb, self.probes = self.b0, -1, False, 0
    self.view = prefetched          # first-bucket view taken before ANY agent ran (the up-front loads)
    self.state = "PROCESS" if prefetched is not None else "LOAD"
    self.result = None              # ("found"|"new", slot)

  def _snapshot(self, b):
    return list(self.m.keys[b * 8:(b + 1) * 8])

  def step(self):
    m = self.m
    if self.state == "LOAD":
      self.view = self._snapshot(self.b)
      self.state = "PROCESS"
    elif self.state == "PROCESS":
      v = self.view
      if self.key in v:
        self.result = ("found", self.b * 8 + v.index(self.key))
        self.state = "DONE"
        return
      free = [i for i, k in enumerate(v) if k in (EMPTY, TOMB)]
      if self.first_free < 0 and free:
        self.first_free = self.b * 8 + free[0]
        self.ff_tomb = v[free[0]] == TOMB
      self.probes += 1
      if EMPTY in v or self.probes >= m.nb:
        assert self.first_free >= 0, "table full"
        self.state = "CAS"
      else:
        self.b = (self.b + 1) % m.nb
        self.state = "LOAD"
    elif self.state == "CAS":
      expect = TOMB if self.ff_tomb else EMPTY
      old = m.keys[self.first_free]
      if old == expect:
        m.keys[self.first_free] = self.key      # atomicCAS succeeded
        self.result = ("new", self.first_free)
        self.state = "DONE"
      elif old == self.key:
        self.result = ("found", self.first_free)
        self.state = "DONE"
      else:                                      # slot taken by another key: rescan with fresh loads
        self.b, self.first_free, self.ff_tomb, self.probes = self.b0, -1, False, 0
        self.state = "LOAD"


def _interleave(seed, nb, n_resident, batch_keys, stale_prefetch):
  rng = np.random.default_rng(seed)
  m = LayoutModel(nb)
  resident = list(rng.choice(np.arange(1000, 1000 + 4 * nb * 8), size=n_resident, replace=False))
  for k in resident:
    m.insert(int(k), 0)
  for k in resident[::3]:                        # leave tombstones / freed slots behind
    m.remove(int(k))
  before = {k for k in m.keys if k not in (EMPTY, TOMB)}
  agents = []
  for k in batch_keys:
    pre = list(m.keys[m.home(k) * 8:(m.home(k) + 1) * 8]) if stale_prefetch else None
    agents.append(_Agent(m, int(k), pre))
  live = list(agents)
  while live:
    a = live[int(rng.integers(len(live)))]
    a.step()
    if a.state == "DONE":
      live.remove(a)
  stored = [k for k in m.keys if k not in (EMPTY, TOMB)]
  assert len(stored) == len(set(stored)), "a key was inserted twice"
  for k in set(int(x) for x in batch_keys):
    news = [a for a in agents if a.key == k and a.result[0] == "new"]
    assert len(news) == (0 if k in before else 1), (k, len(news))
    slots = {a.result[1] for a in agents if a.key == k}
    assert len(slots) == 1 and m.keys[next(iter(slots))] == k     # every duplicate resolved to THE slot of k
  assert set(stored) == before | set(int(x) for x in batch_keys)


@pytest.mark.parametrize("stale_prefetch", [False, True])
def test_find_or_claim_protocol_interleavings(stale_prefetch):
  """Random schedules of many subgroups claiming slots in few buckets, with duplicate keys in the batch, recycled
  tombstones and (stale_prefetch) first-bucket views loaded before any claim: never a duplicate slot, exactly one
  'new' per absent key, and every duplicate of a key ends on the same slot."""
  for seed in range(120):
    rng = np.random.default_rng(10_000 + seed)
    nb = int(rng.integers(2, 5))
    n_res = int(rng.integers(0, nb * 8 * 0.5))
    room = int(nb * 8 * 0.75) - n_res
    n_new = int(rng.integers(1, max(2, room)))
    fresh = rng.choice(np.arange(0, 900), size=n_new, replace=False)
    dup = rng.choice(fresh, size=int(rng.integers(0, n_new + 1)))          # duplicates inside the batch
    batch = np.concatenate([fresh, dup])
    rng.shuffle(batch)
    _interleave(seed, nb, n_res, batch, stale_prefetch)


class _RemoveAgent(object):
  """remove_kernel for one key: probe (read-only scan), re-read the key's bucket, CAS key -> EMPTY | TOMB."""

  def __init__(self, model, key):
    self.m, self.key = model, key
    self.b, self.probes = model.home(key), 0
    self.state, self.slot, self.removed = "SCAN", -1, False

  def step(self):
    m = self.m
    if self.state == "SCAN":
      v = list(m.keys[self.b * 8:(self.b + 1) * 8])
      if self.key in v:
        self.slot = self.b * 8 + v.index(self.key)
        self.state = "REREAD"
      elif EMPTY in v or self.probes + 1 >= m.nb:
        self.state = "DONE"
      else:
        self.probes += 1
        self.b = (self.b + 1) % m.nb
    elif self.state == "REREAD":
      lo = (self.slot // 8) * 8
      self.has_empty = EMPTY in list(m.keys[lo:lo + 8])
      self.state = "CAS"
    elif self.state == "CAS":
      if m.keys[self.slot] == self.key:
        m.keys[self.slot] = EMPTY if self.has_empty else TOMB
        self.removed = True
      self.state = "DONE"


def test_concurrent_removes_keep_every_survivor_reachable():
  for seed in range(150):
    rng = np.random.default_rng(seed)
    nb = int(rng.integers(2, 6))
    m = LayoutModel(nb)
    keys = [int(k) for k in rng.choice(np.arange(0, 5000), size=int(nb * 8 * 0.75), replace=False)]
    for k in keys:
      m.insert(k, k)
    victims = [keys[i] for i in rng.choice(len(keys), size=len(keys) // 2, replace=False)]
    batch = victims + [victims[i] for i in rng.integers(0, len(victims), size=3)] + [7777, 8888]   # dups + absent keys
    agents = [_RemoveAgent(m, k) for k in batch]
    live = list(agents)
    while live:
      a = live[int(rng.integers(len(live)))]
      a.step()
      if a.state == "DONE":
        live.remove(a)
    survivors = set(keys) - set(victims)
    for k in victims:
      assert sum(a.removed for a in agents if a.key == k) == 1      # exactly one duplicate wins the CAS
      assert m.find(k) < 0
    for k in survivors:
      assert m.find(k) >= 0 and m.keys[m.find(k)] == k              # no chain was cut by an erase-to-EMPTY
