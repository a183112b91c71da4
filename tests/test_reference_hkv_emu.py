"""The reference's HKV table tests (kernel_tests/hkv_hashtable_ops_test.py:76-625 and cuckoo_hashtable_ops_test.py:76-99,
202-267), restated with torch tensors over the emulated library: `de.HkvHashTableCreator` with the reference's configs
(the reference's default eviction strategy, LRU, is selected explicitly here -- DESIGN.md 4b).  `test_reach_max_hbm`
(:627-690) is in tests/test_spill_gpu.py / test_mirror_emu.py, the eviction tests in tests/test_evict_gpu.py."""
import numpy as np
import pytest
import torch

from tests.emu import backend


@pytest.fixture(autouse=True)
def _emu():
  with backend.installed():
    yield


def _de():
  from recommenders_addons_b200 import dynamic_embedding as de
  return de


def _hkv(name, value_dtype=torch.int32, dim=8, initializer=0, strategy="LRU", **cfg):
  de = _de()
  cfg.setdefault("evict_strategy", None if strategy is None else de.HkvEvictStrategy[strategy])
  return de.get_variable(name, key_dtype=torch.int64, value_dtype=value_dtype, devices=["cpu"], initializer=initializer,
                         dim=dim, init_size=cfg.get("init_capacity", 1024),
                         kv_creator=de.HkvHashTableCreator(config=de.HkvHashTableConfig(**cfg)))


def test_basic():
  """:76-100"""
  t = _hkv("basic0", init_capacity=1024, max_capacity=99999, reserved_key_start_bit=1)
  assert int(t.size()) == 0
  t.clear()
  assert isinstance(t.tables[0], _de().HkvHashTable) and t.tables[0].evict_strategy == _de().HkvEvictStrategy.LRU


@pytest.mark.parametrize("init_size,expect", [(54321, 54321), (0, 1024 * 1024)])
def test_set_init_size(init_size, expect):
  """:182-217: init_capacity is honoured; 0 means the GPU default of 1024 * 1024 (kDefaultGpuInitCapacity)"""
  de = _de()
  cfg = de.HkvHashTableConfig(init_capacity=init_size, max_capacity=2 * 1024 * 1024, evict_strategy=de.HkvEvictStrategy.LRU)
  table = de.HkvHashTable(torch.int64, torch.int32, torch.zeros(8, dtype=torch.int32), config=cfg, device="cpu",
                          init_size=0 if init_size else 1024 * 1024)
  assert table.capacity() >= expect and table.capacity() < expect + 8


def test_import_and_export():
  """:219-246 (168 keys; the same case as cuckoo_hashtable_ops_test.py:76-99)"""
  t = _hkv("2021-0", dim=3, init_capacity=128, max_capacity=99999)
  keys = torch.arange(168)
  t.upsert(keys, torch.ones(168, 3, dtype=torch.int32))
  assert int(t.size()) == 168
  ek, ev = t.export()
  assert sorted(ek.tolist()) == keys.tolist() and len(ev) == 168


@pytest.mark.parametrize("value_dtype", [torch.float32, torch.int32, torch.int64, torch.int8])
@pytest.mark.parametrize("dim", [1, 10, 64, 200])
def test_insert(value_dtype, dim):
  """:248-290: 18 upserts of 85 fresh keys each, the size follows"""
  t = _hkv("test_insert-%s-%d" % (str(value_dtype).split(".")[-1], dim), value_dtype=value_dtype, dim=dim, initializer=-1,
           init_capacity=102400, max_capacity=102400)
  for i in range(18):
    keys = torch.arange(85 * i, 85 * (i + 1))
    vals = (keys % 100).reshape(-1, 1).repeat(1, dim).to(value_dtype)
    t.upsert(keys, vals)
    assert int(t.size()) == (i + 1) * 85
  got = t.lookup(torch.tensor([0, 99, 1529, 5000]))
  assert got[:, 0].tolist() == [0, 99, 29, -1]
  t.clear()
  assert int(t.size()) == 0


def test_variable_find_with_exists_and_accum():
  """:359-444: the four insert_or_accum cases through the HKV table: {0 -> 10, 2 -> 2, 3 -> 13, 100 -> 99}"""
  dim = 8
  t = _hkv("hkv-accum", value_dtype=torch.float32, dim=dim, initializer=-1.0, init_capacity=1024, max_capacity=1024)
  R = lambda vals: torch.tensor(vals, dtype=torch.float32).reshape(-1, 1).repeat(1, dim)  # noqa: E731
  t.upsert(torch.tensor([0, 1, 2, 3]), R([0, 1, 2, 3]))
  accum_keys = torch.tensor([0, 1, 100, 3])
  old, exists = t.lookup(accum_keys, return_exists=True)
  assert exists.tolist() == [True, True, False, True]
  t.upsert(torch.tensor([100]), R([99]))       # a concurrent writer adds 100 ...
  t.remove(torch.tensor([1]))                  # ... and removes 1 between the find and the accum
  t.accum(accum_keys, R([0, 1, 2, 3]), R([10, 11, 100, 13]), exists)
  assert int(t.size()) == 4
  ek, ev = t.export()
  got = {int(k): float(v[0]) for k, v in zip(ek, ev)}
  assert got == {0: 10.0, 2: 2.0, 3: 13.0, 100: 99.0}


def test_insert_repeat_data():
  """:572-625: the same 50k... (scaled: 20k) keys upserted twice: the size does not change"""
  n = 20000
  t = _hkv("t1-repeat", value_dtype=torch.int64, dim=4, initializer=-1, init_capacity=1 << 16, max_capacity=1 << 16)
  keys = torch.arange(n)
  vals = keys.reshape(-1, 1).repeat(1, 4)
  for i in range(2):
    assert int(t.size()) == (0 if i == 0 else n)
    t.upsert(keys, vals)
    assert int(t.size()) == n


def test_save_and_load_all_with_local_file_system(tmp_path):
  """cuckoo_hashtable_ops_test.py:202-267 / hkv :292-357: every shard of a 2-shard variable saved, then ALL files loaded
  into each table of a fresh variable (load_entire_dir)"""
  de = _de()
  rng = np.random.default_rng(0)
  keys = torch.from_numpy(rng.choice(1 << 30, 10000, replace=False).astype(np.int64))
  vals = torch.from_numpy(rng.integers(-100, 100, (10000, 8)).astype(np.int32))
  src = de.get_variable("t_save_all", value_dtype=torch.int32, dim=8, devices=["cpu"] * 2, initializer=0)
  src.upsert(keys, vals)
  src.save_to_file_system(str(tmp_path), dirpath_env="__unset__", buffer_size=4096)
  dst = de.HkvHashTable(torch.int64, torch.int32, torch.zeros(8, dtype=torch.int32), device="cpu",
                        config=de.HkvHashTableConfig(init_capacity=1 << 15, max_capacity=1 << 15))
  first = src._saved_file_name(0, 1, 0)
  dst.load_from_file_system(str(tmp_path), file_name=first, dirpath_env="__unset__", load_entire_dir=True, buffer_size=4096)
  assert int(dst.size()) == 10000
  o = torch.argsort(keys)
  got, ex = dst.lookup(keys[o], return_exists=True)
  assert bool(ex.all()) and torch.equal(got, vals[o])
