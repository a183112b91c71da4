"""Restrict policies (reference: python/ops/restrict_policies.py; tests kernel_tests/restrict_policies_test.py:130-330)
on a dict-backed table: the policies are pure compositions of table ops, so their logic is checked here without a GPU;
tests/test_restrict_gpu.py runs the same flows on the CUDA tables."""
import numpy as np
import pytest
import torch

from recommenders_addons_b200 import dynamic_embedding as de
from recommenders_addons_b200.dynamic_embedding import restrict_policies as rp
from recommenders_addons_b200.dynamic_embedding import variable as var_mod
from tests.fake_table import DictTableCreator


@pytest.fixture(autouse=True)
def _fresh(monkeypatch):
  var_mod._reset_variables()
  clock = {"t": 1_700_000_000}
  monkeypatch.setattr(rp.time, "time", lambda: clock["t"])
  yield clock
  var_mod._reset_variables()


def make_var(name, policy=None):
  return de.get_variable(name, key_dtype=torch.int64, value_dtype=torch.float32, initializer=-0.1, dim=2,
                         devices=["cpu"], kv_creator=DictTableCreator(), restrict_policy=policy)


def K(a):
  return torch.as_tensor(np.asarray(list(a), dtype=np.int64))


def status_by_key(policy):
  keys, st = policy.status.export()
  return dict(zip(keys.tolist(), st.reshape(-1).tolist()))


def train_step(var, ids):
  """what one optimizer step does to the tables: rows of `ids` are written, the policy sees the ids
  (python/ops/embedding_weights.py:434-444)"""
  var.upsert(ids, var.lookup(ids) - 0.01)
  var.restrict_policy.apply_update(ids)


def test_timestamp_apply_update(_fresh):
  """restrict_policies_test.py:132-165"""
  var = make_var("sp_var_t1")
  policy = de.TimestampRestrictPolicy(var)
  assert int(policy.status.size()) == 0
  policy.apply_update(K(range(3)))
  assert int(policy.status.size()) == 3
  _fresh["t"] += 1
  policy.apply_update(K(range(1, 4)))
  assert int(policy.status.size()) == 4
  st = status_by_key(policy)
  assert all(st[0] < st[y] for y in (1, 2, 3))


def test_timestamp_apply_restriction(_fresh):
  """:167-228"""
  var = make_var("sp_var_t2", de.TimestampRestrictPolicy)
  train_step(var, K(range(6)))
  _fresh["t"] += 1
  train_step(var, K(range(4, 9)))
  all_vars = [var, var.restrict_policy.status]
  assert all(int(v.size()) == 9 for v in all_vars)
  st = status_by_key(var.restrict_policy)
  assert all(st[x] < st[y] for x in range(4) for y in range(4, 9))
  var.restrict_policy.apply_restriction(5, trigger=100)
  assert all(int(v.size()) == 9 for v in all_vars)
  var.restrict(5, trigger=5)
  assert all(int(v.size()) == 5 for v in all_vars)
  assert sorted(var.export()[0].tolist()) == list(range(4, 9))


def test_frequency_apply_update():
  """:233-266"""
  var = make_var("sp_var_f1")
  policy = de.FrequencyRestrictPolicy(var)
  assert int(policy.status.size()) == 0
  policy.apply_update(K(range(3)))
  assert int(policy.status.size()) == 3
  policy.apply_update(K(range(1, 4)))
  assert int(policy.status.size()) == 4
  st = status_by_key(policy)
  assert all(st[x] < st[y] for x in (0, 3) for y in (1, 2))
  assert st == {0: 1, 1: 2, 2: 2, 3: 1}


def test_frequency_apply_restriction():
  """:268-328"""
  var = make_var("sp_var_f2", de.FrequencyRestrictPolicy)
  train_step(var, K(range(6)))
  train_step(var, K(range(4, 9)))
  all_vars = [var, var.restrict_policy.status]
  assert all(int(v.size()) == 9 for v in all_vars)
  st = status_by_key(var.restrict_policy)
  assert all(st[x] < st[y] for x in (0, 1, 2, 3, 6, 7, 8) for y in (4, 5))
  var.restrict_policy.apply_restriction(2, trigger=100)
  assert all(int(v.size()) == 9 for v in all_vars)
  var.restrict_policy.apply_restriction(2, trigger=2)
  assert all(int(v.size()) == 2 for v in all_vars)
  assert sorted(var.export()[0].tolist()) == [4, 5]


def test_duplicate_ids_count_once_per_step_and_argument_errors():
  var = make_var("sp_var_f3", de.FrequencyRestrictPolicy)
  var.restrict_policy.apply_update(K([7, 7, 7, 8]))
  assert status_by_key(var.restrict_policy) == {7: 1, 8: 1}
  with pytest.raises(TypeError):
    var.restrict_policy.apply_restriction(2.5)
  with pytest.raises(TypeError):
    var.restrict_policy.apply_restriction(2, trigger="x")
  with pytest.raises(ValueError):
    var.restrict_policy.apply_restriction(-1)
  with pytest.raises(TypeError):
    make_var("sp_var_f4", policy=object)
  assert make_var("sp_var_f5").restrict(3) is None     # no policy: no-op (dynamic_embedding_variable.py:871-873)
