"""Restrict policies on the CUDA tables with the fused optimizer driving `apply_update`
(reference: kernel_tests/restrict_policies_test.py:167-228, 268-328).  Only table ops validated in round 1 run
underneath.  First hardware run: round 1's driver box (passed on a fresh B200); ungated in round 2."""
import os
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


DEV = "cuda"   # tests/test_mirror_emu.py re-runs this body over the emulated library with DEV = "cpu"


def K(a):
  return torch.as_tensor(np.asarray(list(a), dtype=np.int64), device=DEV)


def status_by_key(policy):
  keys, st = policy.status.export()
  return dict(zip(keys.tolist(), st.reshape(-1).tolist()))


def _train(de, var, opt, ids):
  emb, tw = de.embedding_lookup(var, ids, return_trainable=True)
  loss = (emb * emb).sum()
  loss.backward()
  opt.apply_gradients([(tw.values.grad, tw)])


@pytest.mark.parametrize("policy_name,first,second,overdue,updated", [
    ("TimestampRestrictPolicy", range(6), range(4, 9), range(4), range(4, 9)),
    ("FrequencyRestrictPolicy", range(6), range(4, 9), [0, 1, 2, 3, 6, 7, 8], [4, 5]),
])
def test_apply_restriction_with_fused_adagrad(policy_name, first, second, overdue, updated):
  from recommenders_addons_b200 import dynamic_embedding as de
  var = de.get_variable("sp_var_gpu_" + policy_name, key_dtype=torch.int64, value_dtype=torch.float32, initializer=-0.1,
                        dim=2, num_slot_planes=1, devices=[DEV], restrict_policy=getattr(de, policy_name))
  opt = de.DynamicEmbeddingOptimizer(de.FusedAdagrad(learning_rate=0.1))
  _train(de, var, opt, K(first))
  if policy_name.startswith("Timestamp"):
    time.sleep(1.1)
  _train(de, var, opt, K(second))
  all_vars = [var, var.restrict_policy.status]
  assert all(int(v.size()) == 9 for v in all_vars)
  st = status_by_key(var.restrict_policy)
  assert all(st[x] < st[y] for x in overdue for y in updated)
  var.restrict_policy.apply_restriction(len(list(updated)), trigger=100)
  assert all(int(v.size()) == 9 for v in all_vars)
  var.restrict_policy.apply_restriction(len(list(updated)), trigger=len(list(updated)))
  assert all(int(v.size()) == len(list(updated)) for v in all_vars)
  assert sorted(var.export()[0].tolist()) == list(updated)
  # the optimizer's accumulator rows went with their keys (slot planes are co-indexed with the value rows)
  ks, acc = var.tables[0].export(plane=1)
  assert sorted(ks.tolist()) == list(updated)
