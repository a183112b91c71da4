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
