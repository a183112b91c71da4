"""Sequential model of the score-based capacity management of the GPU table (recommenders_addons_b200/csrc/evict.cu),
built on the layout model (tests/layout_model.py).  It restates the HOST algorithm step by step -- chunking, soft /
hard load limits, exact new-key count, admission against the lowest resident score, selection of the k lowest
scores (threshold + tie quota), erase straight to EMPTY, repair rounds that re-seat keys whose probe chain was
cut -- so that the policy can be property-tested without a GPU (tests/test_evict_model.py) against the
assertions of the reference's own eviction tests
(/root/reference/.../kernel_tests/hkv_hashtable_evict_test.py:110-577).

Score rules = HierarchicalKV v0.1.0-beta.12 as the reference drives it (lookup_table_op_hkv.h:454-547,
python/ops/hkv_hashtable_ops.py:209-216):
  LRU        score = device clock at insert/assign
  LFU        score = old + delta            (delta = provided score, the reference passes ones)
  EPOCHLRU   score = epoch << 32 | low32(clock >> 20)
  EPOCHLFU   score = epoch << 32 | min(low32(old) + delta, 0xffffffff)
  CUSTOMIZED score = provided
A key that is not in the table is admitted only if its score is >= the lowest resident score once the table
is at its limit (HKV refuses a key whose score is below its bucket's minimum)."""
import numpy as np

from tests.layout_model import BUCKET, EMPTY, TOMB, LayoutModel

LRU, LFU, EPOCHLRU, EPOCHLFU, CUSTOMIZED = 0, 1, 2, 3, 4
M32 = 0xffffffff


class EvictModel(LayoutModel):

  def __init__(self, nb, strategy, max_lf=0.875, slab_div=32):
    super().__init__(nb)
    self.strategy = strategy
    self.scores = [0] * (nb * BUCKET + 2)
    self.limit = int(nb * BUCKET * max_lf)
    self.slab = max(1, self.limit // slab_div)
    self.epoch = 0
    self.clock = 1 << 30   # stands for %globaltimer (ns); every mutating launch advances it
    self.n_events = 0
    self.n_refused = 0
    self.n_slow = 0
    self.used_ub = 0     # host upper bound of non-EMPTY slots
    self.snap_used = 0   # last asynchronous snapshot of the device counter

  # ---- score rules ---------------------------------------------------------------------------
  def new_score(self, old, provided):
    s = self.strategy
    if s == LRU:
      return self.clock
    if s == LFU:
      return (old + (1 if provided is None else provided)) & ((1 << 64) - 1)
    if s == EPOCHLRU:
      return (self.epoch << 32) | ((self.clock >> 20) & M32)
    if s == EPOCHLFU:
      return (self.epoch << 32) | min((old & M32) + (1 if provided is None else provided), M32)
    return 0 if provided is None else provided

  # ---- erase keeps the "a free slot has score 0" invariant -------------------------------------
  def remove(self, key):
    s = self.find(key)
    if s >= 0:
      self.scores[s] = 0
    super().remove(key)

  # ---- selection: the k lowest-scored resident (non-special) keys ---------------------------------
  def _live_slots(self):
    return [s for s in range(self.nb * BUCKET) if self.keys[s] not in (EMPTY, TOMB)]

  def select_threshold(self, k):
    """returns (tau, quota): every live score < tau goes, and `quota` of those == tau"""
    sc = sorted(self.scores[s] for s in self._live_slots())
    assert 0 < k <= len(sc)
    tau = sc[k - 1]
    below = sum(1 for x in sc if x < tau)
    return tau, k - below

  def evict_apply(self, tau, quota):
    """one pass over the slots: an evicted slot goes straight back to EMPTY (score 0).  That may cut probe chains
    that ran through its bucket; repair() re-seats the keys that became unreachable.  (The kernel's tie ticket is
    an atomic counter: WHICH of the keys tied at tau go is unspecified; here: lowest slots first.)"""
    evicted = []
    for s in range(self.nb * BUCKET):
      k = self.keys[s]
      if k in (EMPTY, TOMB):
        continue
      sc = self.scores[s]
      go = sc < tau
      if not go and sc == tau and quota > 0:
        quota -= 1
        go = True
      if go:
        evicted.append(k)
        self.keys[s] = EMPTY
        self.scores[s] = 0
        self.vals.pop(s, None)
        self.size -= 1
        self.used -= 1
    return evicted

  def reachable(self, s):
    k = self.keys[s]
    b, bs = self.home(k), s // BUCKET
    while b != bs:
      lo = b * BUCKET
      if any(self.keys[q] == EMPTY for q in range(lo, lo + BUCKET)):
        return False
      b = (b + 1) % self.nb
    return True

  def repair_round(self, order=None):
    """one round = two passes, each changing the set of EMPTY slots in one direction only (the two kernels
    repair_move_kernel / repair_sweep_kernel):
      move : every unreachable key is COPIED to the first free slot of its chain (closer to home); EMPTY slots
             only disappear, so reachable keys -- fresh copies included -- stay reachable during the pass
      sweep: a displaced key that has an earlier match along its chain is a stale copy -> its slot becomes EMPTY
    Returns moves + erasures."""
    moves = erased = 0
    slots = list(range(self.nb * BUCKET)) if order is None else list(order)
    for s in slots:
      k = self.keys[s]
      if k in (EMPTY, TOMB) or self.reachable(s):
        continue
      found, first_free = -1, -1
      for _, chain_slots in self._chain(k):
        for q in chain_slots:
          if found < 0 and self.keys[q] == k:
            found = q
          if first_free < 0 and self.keys[q] in (EMPTY, TOMB):
            first_free = q
      if found >= 0:
        continue          # another copy is reachable already
      assert first_free >= 0
      if self.keys[first_free] == EMPTY:
        self.used += 1
      self.keys[first_free] = k
      self.vals[first_free] = self.vals[s]
      self.scores[first_free] = self.scores[s]
      moves += 1
    for s in slots:
      k = self.keys[s]
      if k in (EMPTY, TOMB) or self.home(k) == s // BUCKET:
        continue
      first = self.find(k)
      if first >= 0 and first != s:
        self.keys[s] = EMPTY
        self.scores[s] = 0
        self.vals.pop(s, None)
        self.used -= 1
        erased += 1
    return moves + erased

  def repair(self):
    rounds = 0
    while self.repair_round():
      rounds += 1
      assert rounds < 64
    self.max_repair_rounds = max(getattr(self, "max_repair_rounds", 0), rounds)

  def purge_tombstones(self):
    """every TOMBSTONE becomes EMPTY, then the repair rounds re-seat the keys whose chains that cut"""
    for s in range(self.nb * BUCKET):
      if self.keys[s] == TOMB:
        self.keys[s] = EMPTY
        self.used -= 1
    self.repair()

  def evict_lowest(self, k):
    live = len(self._live_slots())
    k = min(k, live)
    if k <= 0:
      return []
    tau, quota = self.select_threshold(k)
    ev = self.evict_apply(tau, quota)
    self.repair()
    self.n_events += 1
    return ev

  def _make_room(self, n_adm):
    """bring used + n_adm down to limit - slab (the table cannot grow)"""
    need = self.used + n_adm - self.limit
    if need <= 0:
      return True
    self.evict_lowest(need + self.slab)
    return self.used + n_adm <= self.limit   # false only when tombstones of user removes clog the table

  # ---- mutating entry points --------------------------------------------------------------------
  def room(self, ks, ps, admission):
    """host-side room logic of one launch of len(ks) keys; returns the may-claim mask (None = all)"""
    n = len(ks)
    hard = int(self.nb * BUCKET * 0.95)
    self.used_ub = min(self.used_ub, self.snap_used)   # the snapshot of the previous launch has landed
    if self.used_ub + n <= hard and self.snap_used < self.limit:
      self.used_ub += n
      return None
    # slow path: exact counters (one sync on the GPU)
    self.n_slow += 1
    self.used_ub = self.used
    if self.used + n <= self.limit:
      self.used_ub += n
      return None
    missing = [i for i, k in enumerate(ks) if self.find(k) < 0]
    admit = [True] * n
    n_adm = len(missing)
    if self.used + n_adm > self.limit and admission:
      live = self._live_slots()
      smin = min(self.scores[s] for s in live) if live else 0
      n_adm = 0
      for i in missing:
        admit[i] = self.new_score(0, ps[i]) >= smin
        n_adm += admit[i]
    if self.used > self.size - sum(self.special):
      self.purge_tombstones()          # tombstones of user removes go first (in place), before any key is evicted
    if not self._make_room(n_adm):
      raise RuntimeError("table full")
    self.used_ub = self.used + n_adm
    return admit

  def insert_scored(self, keys, values, scores=None):
    """det_insert on a table with an eviction strategy.  Returns the keys that were refused."""
    refused = []
    chunk = max(1, self.limit // 4)
    for c0 in range(0, len(keys), chunk):
      ks = [int(k) for k in keys[c0:c0 + chunk]]
      vs = values[c0:c0 + chunk]
      ps = [None] * len(ks) if scores is None else [int(x) for x in scores[c0:c0 + chunk]]
      self.clock += 1 << 21
      admit = self.room(ks, ps, True)
      for i, k in enumerate(ks):
        s = self.find(k)
        if s < 0 and admit is not None and not admit[i]:
          refused.append(k)
          continue
        old = self.scores[s] if s >= 0 else 0
        self.insert(k, vs[i])
        s = self.find(k)
        self.scores[s] = self.new_score(old, ps[i])
      self.snap_used = self.used   # the asynchronous snapshot taken after the launch
    self.n_refused += len(refused)
    return refused

  def export_keys_and_scores(self):
    sl = self._live_slots()
    ks = np.array([self.keys[s] for s in sl], dtype=np.int64)
    sc = np.array([self.scores[s] for s in sl], dtype=np.uint64)
    return ks, sc

  def check_invariants(self):
    super().check_invariants()
    for s in range(self.nb * BUCKET):
      if self.keys[s] in (EMPTY, TOMB):
        assert self.scores[s] == 0, "a free slot keeps a stale score"
    assert self.used <= self.nb * BUCKET
