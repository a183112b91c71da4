"""The reference's variable-level tests (kernel_tests/dynamic_embedding_variable_test.py:565-588 and :1380-1830),
restated with torch tensors and run END TO END on the CPU over the emulated libdetable (tests/emu/backend.py).
`test_variable` / `test_variable_find_with_exists_and_accum` (:394-563) are in tests/test_detable_emu.py,
tests/test_oracle.py and -- on the GPU -- tests/test_table_gpu.py; restrict policies (:1831-1957) in
tests/test_restrict_policies.py / test_mirror_emu.py; the checkpoint tests are TF saver plumbing (out of scope)."""
import numpy as np
import pytest
import torch

from tests.emu import backend

I32 = torch.int32


@pytest.fixture(autouse=True)
def _emu():
  with backend.installed():
    yield


def _var(name, value_dtype=I32, initializer=-1, dim=1, **kw):
  from recommenders_addons_b200 import dynamic_embedding as de
  return de.get_variable(name, key_dtype=torch.int64, value_dtype=value_dtype, initializer=initializer, dim=dim,
                         devices=["cpu"], **kw)


def T(x, dtype=torch.int64):
  return torch.tensor(x, dtype=dtype)


def test_variable_initializer():
  """:565-588: 2^17 default rows of dim 10: constant -1 -> mean -1 / std 0; N(0, 0.01) -> mean 0 / std 0.01 (2e-5)"""
  keys = torch.arange(2**17)
  g = torch.Generator().manual_seed(2)
  for i, (init, mean, std) in enumerate([(-1.0, -1.0, 0.0),
                                         (lambda s: torch.randn(list(s), generator=g, dtype=torch.float64) * 0.01, 0.0, 0.01)]):
    vals = _var("t1%d" % i, value_dtype=torch.float32, initializer=init, dim=10).lookup(keys).double()
    assert vals.shape == (2**17, 10)
    np.testing.assert_allclose(vals.mean().item(), mean, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(vals.std(unbiased=False).item(), std, rtol=2e-5, atol=2e-5)


def test_get_variable_reuse_error():
  """:1407-1419"""
  _var("t900", dim=2)
  with pytest.raises(ValueError, match="t900"):
    _var("t900", dim=2)


def test_dynamic_embedding_variable():
  """:1460-1496"""
  table = _var("t10", initializer=T([-1, -2]), dim=2)
  assert int(table.size()) == 0
  table.upsert(T([0, 1, 2, 3]), T([[0, 1], [2, 3], [4, 5], [6, 7]], I32))
  assert int(table.size()) == 4
  table.remove(T([3, 4]))
  assert int(table.size()) == 3
  out = table.lookup(T([0, 1, 4]))
  assert tuple(out.shape) == (3, 2) and out.tolist() == [[0, 1], [2, 3], [-1, -2]]
  ek, ev = table.export()
  assert np.sort(ek.numpy()).tolist() == [0, 1, 2]
  assert np.sort(ev.numpy(), axis=0).tolist() == np.sort([[4, 5], [2, 3], [0, 1]], axis=0).tolist()


def test_dynamic_embedding_variable_export_insert():
  """:1498-1534"""
  t1 = _var("t101", initializer=T([-1, -1]), dim=2)
  t1.upsert(T([0, 1, 2]), T([[0, 1], [2, 3], [4, 5]], I32))
  assert int(t1.size()) == 3
  expected = [[0, 1], [2, 3], [-1, -1]]
  assert t1.lookup(T([0, 1, 3])).tolist() == expected
  ek, ev = t1.export()
  assert ek.numel() == 3 and ev.numel() == 6
  t2 = _var("t102", initializer=T([-1, -1]), dim=2)
  t2.upsert(ek, ev)
  assert int(t2.size()) == 3 and t2.lookup(T([0, 1, 3])).tolist() == expected


def test_dynamic_embedding_variable_invalid_shape():
  """:1536-1570: "Expected shape" for every value tensor that is not keys.shape + [dim]"""
  table = _var("t110", initializer=T([-1, -1]), dim=2)
  keys = T([0, 1, 2])
  for bad in ([0, 1, 2, 3, 4, 5], [[0, 1, 2], [3, 4, 5]], [[0, 1], [2, 3]], [[0], [2], [4]]):
    with pytest.raises(ValueError, match="Expected shape"):
      table.upsert(keys, T(bad, I32))
  table.upsert(keys, T([[0, 1], [2, 3], [4, 5]], I32))
  assert int(table.size()) == 3


def test_rank_handling():
  """:1572-1687 find_high_rank / insert_low_rank / remove_low_rank / insert_high_rank / remove_high_rank"""
  t = _var("t140")
  t.upsert(T([0, 1, 2]), T([[0], [1], [2]], I32))
  out = t.lookup(T([[0, 1], [2, 4]]))
  assert tuple(out.shape) == (2, 2, 1) and out.tolist() == [[[0], [1]], [[2], [-1]]]
  t = _var("t150")
  t.upsert(T([[0, 1], [2, 3]]), T([[[0], [1]], [[2], [3]]], I32))
  assert int(t.size()) == 4 and t.lookup(T([0, 1, 3, 4])).tolist() == [[0], [1], [3], [-1]]
  t.remove(T([1, 4]))
  assert int(t.size()) == 3 and t.lookup(T([0, 1, 3, 4])).tolist() == [[0], [-1], [3], [-1]]
  t = _var("t170", initializer=T([-1, -1, -1], I32), dim=3)
  t.upsert(T([0, 1, 2]), T([[0, 1, 2], [2, 3, 4], [4, 5, 6]], I32))
  out = t.lookup(T([[0, 1], [3, 4]]))
  assert tuple(out.shape) == (2, 2, 3)
  assert out.tolist() == [[[0, 1, 2], [2, 3, 4]], [[-1, -1, -1], [-1, -1, -1]]]
  t.remove(T([[0, 3]]))
  assert int(t.size()) == 2
  assert t.lookup(T([[0, 1], [2, 3]])).tolist() == [[[-1, -1, -1], [2, 3, 4]], [[4, 5, 6], [-1, -1, -1]]]


def test_several_variables_and_tensor_default():
  """:1689-1744"""
  tabs = [_var("t19%d" % i) for i in range(1, 4)] + [_var("t200", initializer=T(-1, I32))]
  for t in tabs:
    t.upsert(T([0, 1, 2]), T([[0], [1], [2]], I32))
  for t in tabs:
    assert int(t.size()) == 3 and t.lookup(T([0, 1, 3])).tolist() == [[0], [1], [-1]]


def test_signature_mismatch():
  """:1746-1788: keys / values of the wrong dtype raise (ValueError in the reference's Python wrappers)"""
  table = _var("t210")
  keys, values = T([0, 1, 2]), T([[0], [1], [2]], I32)
  with pytest.raises(ValueError):
    table.upsert(torch.tensor([4.0, 5.0, 6.0]), values)
  with pytest.raises(ValueError):
    table.upsert(keys, torch.tensor([[0.5], [1.5], [2.5]]))
  assert int(table.size()) == 0
  table.upsert(keys, values)
  assert int(table.size()) == 3
  with pytest.raises(ValueError):
    table.lookup(T([1, 2, 3], I32))
  with pytest.raises(TypeError):      # the kernels' convention (MatchSignature) is caught as well
    table.lookup(T([1, 2, 3], I32))


def test_int_float_and_random_init():
  """:1790-1829"""
  t = _var("t220", value_dtype=torch.float32, initializer=-1.0)
  t.upsert(T([3, 7, 0]), torch.tensor([[7.5], [-1.2], [9.9]]))
  np.testing.assert_allclose(t.lookup(T([7, 0, 11])).numpy(), [[-1.2], [9.9], [-1.0]], rtol=1e-6)
  g = torch.Generator().manual_seed(1)
  t = _var("t230", value_dtype=torch.float32, initializer=lambda s: torch.rand(list(s), generator=g))
  t.upsert(T([0, 1, 2]), torch.tensor([[0.0], [1.0], [2.0]]))
  res = t.lookup(T([0, 1, 3]))
  assert res[:2].tolist() == [[0.0], [1.0]] and res[2].item() != -1.0 and 0.0 <= res[2].item() < 1.0


def test_verify_embedding_weights_is_a_variable_method():
  """EmbeddingWeights.verify_embedding_weights / verify_embedding_param_weights (embedding_weights.py:53, 78-95;
  Variable's implementation dynamic_embedding_variable.py:694-696): key dtype vs ids, value dtype vs weights"""
  from recommenders_addons_b200 import dynamic_embedding as de
  var = _var("verify-1", value_dtype=torch.float32, initializer=0.0, dim=4)
  ids = de.SparseIds(torch.tensor([[0, 0], [1, 0]]), T([3, 4]), (2, 1))
  w = de.SparseIds(ids.indices, torch.tensor([1.0, 2.0]), ids.dense_shape)
  var.verify_embedding_weights(ids, w)
  de.Variable.verify_embedding_param_weights(var, ids)
  with pytest.raises(TypeError, match="key_dtype should be same with sparse_ids.dtype"):
    var.verify_embedding_weights(de.SparseIds(ids.indices, T([3, 4], I32), ids.dense_shape))
  with pytest.raises(TypeError, match="value_dtype should be same with sparse_weights.dtype"):
    var.verify_embedding_weights(ids, de.SparseIds(ids.indices, torch.tensor([1.0, 2.0], dtype=torch.float64), ids.dense_shape))
  with pytest.raises(ValueError, match="Missing embedding_weights"):
    de.Variable.verify_embedding_param_weights(None, ids)
