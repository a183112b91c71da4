"""Capacity management on the GPU (csrc/evict.cu) through the reference-facing API: the reference's own eviction
tests (kernel_tests/hkv_hashtable_evict_test.py:110-577) with torch tensors, plus engine-level checks (explicit
det_evict against a sort, survivors keep their rows, scores follow their keys through growth, accum / fused
optimizer refresh scores, removed keys leave score 0).  The same bodies run on the sequential model in
tests/test_evict_model.py.

First hardware run: round 1's driver box (all five suites passed on a fresh B200); ungated in round 2."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DIM = 8
DEV = "cuda"   # tests/test_mirror_emu.py re-runs these bodies over the emulated library with DEV = "cpu"
CHURN = (1 << 15, 120, 1500)   # steady-state test: (capacity, steps, new keys per step); smaller on the emulator


def _de():
  from recommenders_addons_b200 import dynamic_embedding as de
  return de


def make_table(strategy, name, capacity=1024, init_capacity=None, step_per_epoch=0, gen_scores_fn=None,
               value_dtype=torch.int32, dim=DIM, num_slot_planes=0):
  de = _de()
  return de.get_variable(name, key_dtype=torch.int64, value_dtype=value_dtype, initializer=0, dim=dim, devices=[DEV],
                         init_size=capacity, num_slot_planes=num_slot_planes,
                         kv_creator=de.HkvHashTableCreator(config=de.HkvHashTableConfig(
                             init_capacity=init_capacity or capacity, max_capacity=capacity,
                             evict_strategy=strategy, step_per_epoch=step_per_epoch,
                             gen_scores_fn=gen_scores_fn)))


def K(a):
  return torch.as_tensor(np.asarray(a, dtype=np.int64), device=DEV)


def V(vals, dtype=torch.int32, dim=DIM):
  return torch.as_tensor(np.repeat(np.asarray(vals).reshape(-1, 1), dim, axis=1), device=DEV).to(dtype)


def gen_scores_fn(keys):
  return keys + 1


def export_ks(t):
  ks, sc = t.export_keys_and_scores(1)
  return ks.cpu().numpy(), sc.cpu().numpy()


def check_table(t):
  """no key stored twice, size() == number of exported keys, every exported key is found"""
  ks, vs = t.export()
  ks = ks.cpu().numpy()
  assert len(np.unique(ks)) == len(ks)
  assert int(t.size()) == len(ks)
  if len(ks):
    _, ex = t.lookup(K(ks), return_exists=True)
    assert bool(ex.all())
  for tab in t.tables:
    st = tab.stats()
    assert st["error_flags"] == 0
    assert st["used_slots"] == st["size"]  # eviction leaves no tombstones


@pytest.mark.parametrize("strategy", [0, 1, 2, 3, 4])
def test_evict_strategy_basic_and_export_scores(strategy):
  de = _de()
  t = make_table(de.HkvEvictStrategy(strategy), "ev_basic_%d" % strategy, step_per_epoch=4, gen_scores_fn=gen_scores_fn)
  keys = K([0, 1, 2, 3])
  t.upsert(keys, V([0, 1, 2, 3]))
  assert torch.equal(t.lookup(keys), V([0, 1, 2, 3]))
  ek, es = export_ks(t)
  assert (np.sort(ek) == np.arange(4)).all()
  if strategy == 4:
    assert (np.sort(es) == np.arange(4) + 1).all()
  elif strategy in (1, 3):
    assert (es == 1).all()
  ek2, ev2, es2 = t.export_with_scores(1)
  assert torch.equal(t.lookup(ek2), ev2)
  assert (np.sort(es2.cpu().numpy()) == np.sort(es)).all()
  check_table(t)


def test_evict_strategy_lfu():
  de = _de()
  t = make_table(de.HkvEvictStrategy.LFU, "ev_lfu")
  keys = K([0, 1, 2, 3])
  t.upsert(keys, V([0, 1, 2, 3]))
  assert (export_ks(t)[1] == 1).all()
  t.upsert(keys, V([0, 1, 2, 3]))
  assert (export_ks(t)[1] == 2).all()
  t.upsert(K([0, 1, 4, 5]), V([0, 1, 2, 3]))
  assert (np.sort(export_ks(t)[1]) == np.array([1, 1, 2, 2, 3, 3])).all()
  keys = np.arange(4, 1034)
  t.upsert(K(keys), V([10] * len(keys)))
  ek, es = export_ks(t)
  assert len(ek) < 1024
  assert (np.sort(ek)[:6] == np.arange(0, 6)).all()
  assert (np.sort(es)[-6:] == np.array([2, 2, 2, 2, 3, 3])).all()
  check_table(t)


def test_evict_strategy_epoch_lfu():
  de = _de()
  t = make_table(de.HkvEvictStrategy.EPOCHLFU, "ev_elfu", step_per_epoch=4)
  for base in [1, 1 + (1 << 32), 1 + (2 << 32)]:
    keys = K([0, 1, 2, 3])
    t.upsert(keys, V([0, 1, 2, 3]))
    assert (np.sort(export_ks(t)[1])[-4:] >= base).all()
    t.upsert(keys, V([0, 1, 2, 3]))
    assert (np.sort(export_ks(t)[1])[-4:] >= base + 1).all()
    t.upsert(K([0, 1, 4, 5]), V([0, 1, 2, 3]))
    assert (np.sort(export_ks(t)[1])[-6:] >= np.array([base, base, base + 1, base + 1, base + 2, base + 2])).all()
    keys = np.arange(4, 1024)
    t.upsert(K(keys), V([10] * len(keys)))
    ek, es = export_ks(t)
    assert len(ek) < 1024
    assert (np.sort(ek)[:6] == np.arange(0, 6)).all()
    assert (np.sort(es)[-6:] >= np.array([base + 1] * 4 + [base + 2] * 2)).all()
  check_table(t)


def test_evict_strategy_lru():
  de = _de()
  t = make_table(de.HkvEvictStrategy.LRU, "ev_lru")
  keys = K([0, 1, 2, 3])
  t.upsert(keys, V([0, 1, 2, 3]))
  assert np.isin(np.arange(4), export_ks(t)[0]).all()
  t.upsert(K([2, 3, 6, 7]), V([0, 1, 2, 3]))
  ek, es = export_ks(t)
  l1 = [int(s) for k, s in zip(ek, es) if k in (0, 1)]
  l2 = [int(s) for k, s in zip(ek, es) if k in (2, 3)]
  assert max(l1) < min(l2)
  keys = np.arange(4, 1044)
  t.upsert(K(keys), V([10] * len(keys)))
  keys = np.arange(1024, 1400)
  t.upsert(K(keys), V([10] * len(keys)))
  ek, _ = export_ks(t)
  assert len(ek) <= 1024
  assert not np.isin(np.arange(0, 4), ek).any()
  check_table(t)


def test_evict_strategy_epoch_lru():
  de = _de()
  t = make_table(de.HkvEvictStrategy.EPOCHLRU, "ev_elru", step_per_epoch=1)
  for epoch in range(2):
    keys = np.arange(0, 1024)
    t.upsert(K(keys), V([10] * len(keys)))
    _, es = export_ks(t)
    assert (es >= (epoch << 32)).all()
    assert (es < (epoch << 32) + 0xffffffff).all()
  check_table(t)


def test_evict_strategy_custom():
  de = _de()
  calls = [0]

  def gen_custom(keys):
    calls[0] += 1
    return torch.full((keys.numel(),), 10000 if calls[0] == 1 else 1, dtype=torch.int64, device=keys.device)

  t = make_table(de.HkvEvictStrategy.CUSTOMIZED, "ev_custom", gen_scores_fn=gen_custom)
  keys = np.arange(2048, 4096)
  t.upsert(K(keys), V([10] * len(keys)))
  keys = np.arange(0, 1024)
  t.upsert(K(keys), V([10] * len(keys)))
  ek, es = export_ks(t)
  assert len(ek) > 0
  assert (es == 10000).all()
  assert (ek >= 1024).all()
  check_table(t)


# ---- engine-level checks -------------------------------------------------------------------------------------
def test_explicit_evict_takes_the_lowest_scores_and_survivors_keep_their_rows():
  de = _de()
  rng = np.random.default_rng(3)
  n = 50000
  t = make_table(de.HkvEvictStrategy.CUSTOMIZED, "ev_explicit", capacity=1 << 17, gen_scores_fn=lambda k: (k * 7919) % 1000003,
                 value_dtype=torch.float32, dim=16)
  keys = rng.choice(1 << 40, size=n, replace=False).astype(np.int64)
  vals = torch.as_tensor(rng.standard_normal((n, 16)).astype(np.float32), device=DEV)
  t.upsert(K(keys), vals)
  scores = (keys * 7919) % 1000003
  assert len(np.unique(scores)) > 0.9 * n
  tab = t.tables[0]
  for k_ev in (1, 777, 20000):
    order = np.argsort(scores, kind="stable")
    kth = scores[order[k_ev - 1]]
    got = tab.evict(k_ev)
    assert got == k_ev
    out, ex = t.lookup(K(keys), return_exists=True)
    ex = ex.cpu().numpy()
    # everything strictly below the k-th score went, everything above stayed; exactly k_ev keys went
    assert not ex[scores < kth].any()
    assert ex[scores > kth].all()
    assert (~ex).sum() == k_ev
    assert torch.equal(out[torch.as_tensor(ex, device=DEV)], vals[torch.as_tensor(ex, device=DEV)])
    keep = ex
    keys, scores, vals = keys[keep], scores[keep], vals[torch.as_tensor(keep, device=DEV)]
    check_table(t)
  assert tab.stats()["evict_events"] == 3 and tab.stats()["evicted_keys"] == 1 + 777 + 20000


def test_scores_follow_their_keys_through_growth():
  de = _de()
  t = make_table(de.HkvEvictStrategy.CUSTOMIZED, "ev_grow", capacity=1 << 16, init_capacity=1024,
                 gen_scores_fn=lambda k: k * 3 + 1)
  keys = np.arange(1, 30001)
  for c in range(0, len(keys), 5000):
    t.upsert(K(keys[c:c + 5000]), V(keys[c:c + 5000] % 1000))
  assert t.tables[0].stats()["rehash_count"] >= 1
  ek, es = export_ks(t)
  assert len(ek) == len(keys)
  assert (es == ek * 3 + 1).all()
  check_table(t)


def test_accum_and_fused_optimizer_refresh_scores_and_remove_clears_them():
  de = _de()
  t = make_table(de.HkvEvictStrategy.LFU, "ev_touch", capacity=4096, value_dtype=torch.float32, dim=8, num_slot_planes=1)
  tab = t.tables[0]
  keys = K(np.arange(100))
  t.upsert(keys, torch.ones(100, 8, device=DEV))
  assert (export_ks(t)[1] == 1).all()
  # accum on resident keys (exists = True) and on new keys (exists = False)
  ks2 = K(np.arange(50, 150))
  exists = torch.as_tensor(np.arange(50, 150) < 100, device=DEV)
  tab.accum(ks2, torch.ones(100, 8, device=DEV), exists)
  ek, es = export_ks(t)
  sc = dict(zip(ek.tolist(), es.tolist()))
  assert all(sc[k] == 1 for k in range(50)) and all(sc[k] == 2 for k in range(50, 100))
  assert all(sc[k] == 1 for k in range(100, 150))
  # one fused Adagrad step touches its keys (the optimizer's update_op is an upsert in the reference)
  opt = de.FusedAdagrad(learning_rate=0.1)
  opt.apply_gradients([(torch.ones(10, 8, device=DEV), (t, K(np.arange(10))))])
  ek, es = export_ks(t)
  sc = dict(zip(ek.tolist(), es.tolist()))
  assert all(sc[k] == 2 for k in range(10)) and all(sc[k] == 1 for k in range(10, 50))
  # removed keys leave score 0 behind: a re-inserted key starts counting from scratch
  t.remove(K(np.arange(50, 100)))
  t.upsert(K(np.arange(50, 100)), torch.ones(50, 8, device=DEV))
  ek, es = export_ks(t)
  sc = dict(zip(ek.tolist(), es.tolist()))
  assert all(sc[k] == 1 for k in range(50, 100))
  check_table(t)


def test_steady_state_churn_keeps_the_table_consistent():
  """many steps at the limit: mixed resident / new keys, LRU; content is checked against the rows last written"""
  de = _de()
  rng = np.random.default_rng(11)
  cap, steps, per = CHURN
  t = make_table(de.HkvEvictStrategy.LRU, "ev_churn", capacity=cap, value_dtype=torch.float32, dim=16)
  written = {}
  nxt = 0
  for step in range(steps):
    new = np.arange(nxt, nxt + per)
    nxt += per
    ek = t.export()[0].cpu().numpy()
    ek = ek[ek >= per]   # the keys of step 0 are never written again: they must be the first to go
    old = rng.choice(ek, size=min(len(ek), per), replace=False) if len(ek) else np.empty(0, dtype=np.int64)
    ks = np.concatenate([new, old]).astype(np.int64)
    vals = rng.standard_normal((len(ks), 16)).astype(np.float32)
    t.upsert(K(ks), torch.as_tensor(vals, device=DEV))
    for k, v in zip(ks.tolist(), vals):
      written[k] = v
  check_table(t)
  ks, vs = t.export()
  ks, vs = ks.cpu().numpy(), vs.cpu().numpy()
  assert 0.8 * 0.875 * cap <= len(ks) <= cap
  for k, v in zip(ks.tolist(), vs):
    assert np.array_equal(written[k], v)
  st = t.tables[0].stats()
  assert st["evict_events"] >= 1
  # the newest keys are resident, the oldest are gone
  _, ex = t.lookup(K(np.arange(nxt - per, nxt)), return_exists=True)
  assert bool(ex.all())
  _, ex = t.lookup(K(np.arange(0, per)), return_exists=True)
  assert not bool(ex.any())
