"""The Python mirror (de.get_variable / de.HkvHashTable / de.Variable / fused optimizers / restrict policies) over the
EMULATED libdetable with CPU tensors (tests/emu/backend.py): the bodies of the GPU suites that have not run on real
hardware yet (tests/test_evict_gpu.py, tests/test_restrict_gpu.py) are executed here as they are, so their Python glue
and the C ABI underneath have been exercised end to end before the first GPU run."""
import os
import time

import pytest

from tests import test_callers_gpu as CG
from tests import test_evict_gpu as EG
from tests import test_restrict_gpu as RG
from tests import test_spill_gpu as SG
from tests.emu import backend


@pytest.fixture(autouse=True)
def _emu(monkeypatch):
  with backend.installed():
    monkeypatch.setattr(EG, "DEV", "cpu")
    monkeypatch.setattr(RG, "DEV", "cpu")
    monkeypatch.setattr(SG, "DEV", "cpu")
    monkeypatch.setattr(CG, "DEV", "cpu")
    monkeypatch.setattr(SG, "REACH", 1 << 14)
    monkeypatch.setattr(SG, "SPLIT", (1 << 15, 64))
    monkeypatch.setattr(RG.time, "sleep", lambda s: _advance(monkeypatch, s))
    yield


_OFFSET = [0.0]
_REAL_TIME = time.time


def _advance(monkeypatch, s):
  """the timestamp policy has one-second resolution: advance its clock instead of sleeping"""
  from recommenders_addons_b200.dynamic_embedding import restrict_policies as rp
  _OFFSET[0] += s
  monkeypatch.setattr(rp.time, "time", lambda: _REAL_TIME() + _OFFSET[0])


@pytest.mark.parametrize("strategy", [0, 1, 2, 3, 4])
def test_evict_strategy_basic_and_export_scores(strategy):
  EG.test_evict_strategy_basic_and_export_scores(strategy)


@pytest.mark.parametrize("name", [
    "test_evict_strategy_lfu", "test_evict_strategy_epoch_lfu", "test_evict_strategy_lru", "test_evict_strategy_epoch_lru",
    "test_evict_strategy_custom", "test_scores_follow_their_keys_through_growth",
    "test_accum_and_fused_optimizer_refresh_scores_and_remove_clears_them",
])
def test_evict_suite_body(name):
  getattr(EG, name)()


def test_explicit_evict(monkeypatch):
  EG.test_explicit_evict_takes_the_lowest_scores_and_survivors_keep_their_rows()


@pytest.mark.parametrize("policy_name,first,second,overdue,updated", [
    ("TimestampRestrictPolicy", range(6), range(4, 9), range(4), range(4, 9)),
    ("FrequencyRestrictPolicy", range(6), range(4, 9), [0, 1, 2, 3, 6, 7, 8], [4, 5]),
])
def test_restrict_policies_with_fused_adagrad(policy_name, first, second, overdue, updated):
  RG.test_apply_restriction_with_fused_adagrad(policy_name, first, second, overdue, updated)


def test_steady_state_churn(monkeypatch):
  monkeypatch.setattr(EG, "CHURN", (1 << 11, 40, 150))
  EG.test_steady_state_churn_keeps_the_table_consistent()


def test_field_wise_embedding_docstring_example():
  """keras/layers/embedding.py:365-395 (the docstring example), forward + one training step, on the emulated library"""
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  nslots = 3
  layer = de.layers.FieldWiseEmbedding(2, nslots, slot_map_fn=lambda ids: ids % nslots, initializer=0.0, devices=["cpu"],
                                       name="fieldwise", num_slot_planes=1)
  ids = torch.tensor([[23, 12, 0], [9, 13, 10]], dtype=torch.int64)
  assert torch.equal(layer(ids), torch.zeros(2, 3, 2))
  layer.params.upsert(torch.arange(100, dtype=torch.int64), torch.ones(100, 2))
  layer.train()
  out = layer(ids)
  exp = torch.tensor([[[2., 2.], [0., 0.], [1., 1.]], [[1., 1.], [2., 2.], [0., 0.]]])
  assert torch.equal(out, exp)
  out.sum().backward()
  layer.apply_gradients(de.FusedAdagrad(0.5, initial_accumulator_value=1.0))
  got = layer.params.lookup(torch.tensor([23, 12, 1], dtype=torch.int64))
  step = 0.5 * 1.0 / (1.0 + 1.0) ** 0.5                    # a += g*g; p -= lr*g/sqrt(a) with g = 1
  assert torch.allclose(got, torch.tensor([[1 - step] * 2, [1 - step] * 2, [1.0, 1.0]]))
  mean = de.layers.FieldWiseEmbedding(2, nslots, slot_map_fn=lambda i: i % nslots, combiner="mean", initializer=1.0,
                                      devices=["cpu"], name="fieldwise-mean")
  assert torch.equal(mean(ids), torch.tensor([[[1., 1.], [0., 0.], [1., 1.]], [[1., 1.], [1., 1.], [0., 0.]]]))
  with pytest.raises(ValueError):
    de.layers.FieldWiseEmbedding(2, 3, slot_map_fn=None, devices=["cpu"], name="bad")
  assert de.layers.BasicEmbedding is de.layers.Embedding


@pytest.mark.parametrize("name", ["test_reach_max_hbm", "test_split_value_plane_all_ops_against_a_dict",
                                  "test_growth_across_the_budget_and_negative_budget"])
def test_spill_suite_body(name):
  """max_hbm_for_vectors: option plumbing, split accounting and every op on a table created with a budget"""
  getattr(SG, name)()


@pytest.mark.parametrize("name", sorted(CG.OPTIMIZERS))
def test_composed_optimizer_matches_dense_twin(name):
  CG.test_composed_optimizer_matches_dense_twin(name)


@pytest.mark.parametrize("name", ["momentum", "rmsprop", "adam", "adagrad"])
def test_composed_optimizer_bp_v2_matches_dense_twin(name):
  CG.test_composed_optimizer_bp_v2_matches_dense_twin(name)


@pytest.mark.parametrize("kind", ["adagrad", "adam"])
def test_composed_and_fused_paths_agree(kind):
  CG.test_composed_and_fused_paths_agree(kind)


@pytest.mark.parametrize("name", ["test_untouched_rows_and_slot_tables_stay_sparse",
                                  "test_get_slot_variables_and_restrict_policy_tracks_them",
                                  "test_layers_and_apply_sparse_take_any_optimizer",
                                  "test_model_mode_and_trainable_wrapper_filter",
                                  "test_shadow_variable_training_matches_dense_twin"])
def test_callers_suite_body(name):
  getattr(CG, name)()


@pytest.mark.parametrize("bp_v2", [False, True])
def test_trainable_wrapper_prefetch_and_update_op(bp_v2):
  CG.test_trainable_wrapper_prefetch_and_update_op(bp_v2)


def test_table_file_ops_append_and_load_entire_dir(tmp_path):
  """attrs of TFRA>...SaveToFileSystem / LoadFromFileSystem (cuckoo_hashtable_ops.cc:257-291): append_to_file,
  load_entire_dir (every `<name>_mht_*` pair of the directory, cuckoo_hashtable_op.cc:477-498), buffer_size"""
  import numpy as np
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  d = str(tmp_path)
  mk = lambda name: de.CuckooHashTable(torch.int64, torch.float32, torch.zeros(4), name=name, device="cpu")  # noqa: E731
  k1, k2 = torch.arange(0, 700), torch.arange(1000, 1500)
  v1, v2 = torch.arange(700.).reshape(-1, 1).repeat(1, 4), -torch.arange(500.).reshape(-1, 1).repeat(1, 4)
  a, b, c = mk("a"), mk("b"), mk("c")
  a.insert(k1, v1)
  b.insert(k2, v2)
  a.save_to_file_system(d, file_name="emb_mht_1of2", buffer_size=256)
  b.save_to_file_system(d, file_name="emb_mht_2of2", buffer_size=256)
  c.insert(torch.tensor([77777]), torch.ones(1, 4))
  c.load_from_file_system(d, file_name="emb_mht_1of2", load_entire_dir=True, buffer_size=300)
  assert int(c.size()) == 1200                                 # cleared once, both shards loaded
  got, ex = c.lookup(torch.cat([k1, k2]), return_exists=True)
  assert bool(ex.all()) and torch.equal(got, torch.cat([v1, v2]))
  c.load_from_file_system(d, file_name="emb_mht_2of2")        # one file: clear + insert
  assert int(c.size()) == 500
  b.save_to_file_system(d, file_name="emb_mht_1of2", append_to_file=True)
  assert len(np.fromfile(os.path.join(d, "emb_mht_1of2-keys"), dtype="<i8")) == 1200
  with pytest.raises(Exception):
    c.load_from_file_system(d, file_name="nothing_mht_1of1", load_entire_dir=True)


def test_smoke_body_on_the_emulated_library():
  """__graft_entry__.smoke() -- what the driver runs on cuda:0 at round end -- with CPU tensors over the emulated
  library: its Python glue and every entry point it calls keep working between GPU runs"""
  import inspect
  import __graft_entry__ as entry
  src = inspect.getsource(entry.smoke)
  for cuda_only in ('assert torch.cuda.is_available(), "smoke() needs cuda:0"', "torch.cuda.set_device(0)",
                    "torch.cuda.synchronize()"):
    assert cuda_only in src
    src = src.replace(cuda_only, "pass")
  make = 'var = de.Variable(dim=dim, initializer=0.5, num_slot_planes=1, name="smoke")'
  assert make in src
  src = src.replace(make, make[:-1] + ', devices=["cpu"])')
  ns = {}
  exec(src, ns)  # pylint: disable=exec-used
  ns["smoke"]()


def test_lookup_sparse_max_norm_fused_vs_composed():
  CG.test_lookup_sparse_max_norm_fused_vs_composed(8)


def test_file_system_saver_reshards_on_restore(tmp_path):
  """de.FileSystemSaver (python/ops/dynamic_embedding_creator.py:415-560): 3 shards saved by one process, restored into
  2 shards of each of 2 processes -- every key lands on the process / shard the partitioner assigns it to"""
  import numpy as np
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  rng = np.random.default_rng(0)
  keys = torch.from_numpy(rng.choice(1 << 40, 5000, replace=False).astype(np.int64))
  vals = torch.from_numpy(rng.normal(size=(5000, 4)).astype(np.float32))
  src = de.get_variable("emb/saver_var", dim=4, devices=["cpu"] * 3, kv_creator=de.CuckooHashTableCreator(
      saver=de.FileSystemSaver()))
  src.upsert(keys, vals)
  de.FileSystemSaver(save_path=str(tmp_path), buffer_size=700).save(src)
  assert len(os.listdir(tmp_path)) == 6
  total = 0
  for rank in range(2):
    from recommenders_addons_b200.dynamic_embedding import variable as V
    V._VARIABLES.pop("emb/saver_var", None)
    dst = de.get_variable("emb/saver_var", dim=4, devices=["cpu"] * 2)
    de.FileSystemSaver(proc_size=2, proc_rank=rank, save_path=str(tmp_path), buffer_size=512).restore(dst)
    mine = de.default_partition_fn(keys, 2) == rank
    assert int(dst.size()) == int(mine.sum())
    got, ex = dst.lookup(keys[mine], return_exists=True)
    assert bool(ex.all()) and torch.equal(got, vals[mine])
    total += int(dst.size())
  assert total == 5000
  with pytest.raises(TypeError):
    de.FileSystemSaverConfig(proc_size=2)
  with pytest.raises(RuntimeError):
    de.CuckooHashTableCreator(saver=object())


@pytest.mark.parametrize("kind", ["adagrad", "adam"])
def test_fused_optimizer_state_survives_a_checkpoint(kind, tmp_path):
  CG.test_fused_optimizer_state_survives_a_checkpoint(kind, tmp_path)


def test_hkv_table_takes_the_capacity_attributes_directly_and_short_file_name_names_slots():
  """hkv_hashtable_ops.py:66-136 (init_capacity / max_capacity / evict_strategy ... as constructor arguments, a config
  overriding them) and `short_file_name` (dynamic_embedding_variable.py:553-562, dynamic_embedding_optimizer.py:882-885)"""
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  t = de.HkvHashTable(torch.int64, torch.float32, [0.0] * 4, name="hkv-direct", init_capacity=256, max_capacity=256,
                      evict_strategy=de.HkvEvictStrategy.LFU, device="cpu")
  assert t.evict_strategy == de.HkvEvictStrategy.LFU and int(t.capacity()) == 256
  t.insert(torch.arange(10), torch.ones(10, 4))
  assert int(t.size()) == 10
  cfg = de.HkvHashTableConfig(init_capacity=512, max_capacity=512, evict_strategy=de.HkvEvictStrategy.LRU)
  t2 = de.HkvHashTable(torch.int64, torch.float32, [0.0] * 4, name="hkv-config-wins", init_capacity=64, max_capacity=64,
                       evict_strategy=de.HkvEvictStrategy.LFU, config=cfg, device="cpu")
  assert t2.evict_strategy == de.HkvEvictStrategy.LRU and int(t2.capacity()) == 512
  var = de.get_variable("sfn", dim=4, initializer=0.0, devices=["cpu"], num_slot_planes=1, short_file_name=True)
  assert [s.name for s in var.get_slot_variables(de.FusedAdagrad(0.1))] == ["sfn/accumulator"]
  var2 = de.get_variable("lfn", dim=4, initializer=0.0, devices=["cpu"], num_slot_planes=1)
  assert [s.name for s in var2.get_slot_variables(de.FusedAdagrad(0.1))] == ["lfn/Adagrad/accumulator"]


def test_file_system_saver_checkpoints_optimizer_state_across_a_reshard(tmp_path):
  CG.test_file_system_saver_checkpoints_optimizer_state_across_a_reshard(tmp_path)


def test_restrict_shrinks_the_fused_optimizer_slots_with_the_variable():
  CG.test_restrict_shrinks_the_fused_optimizer_slots_with_the_variable()


def test_shadow_variable_assign_and_verify():
  """shadow_embedding_ops.py:166-168, :198-226 (kernel_tests/shadow_embedding_ops_test.py test_create / test_read_value):
  assign replaces the lookup buffer of the current ids, update_op writes it back to the table"""
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  var = de.get_variable("shadow-assign", dim=3, initializer=0.0, devices=["cpu"])
  sh = de.shadow_ops.ShadowVariable(var, name="sv-assign")
  out = de.shadow_ops.embedding_lookup(sh, torch.tensor([5, 9]))
  assert out.shape == (2, 3) and int(var.size()) == 0
  got = sh.assign(torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]))
  assert torch.equal(got.detach(), torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])) and torch.equal(sh.value().detach(), got.detach())
  sh.update_op()
  assert torch.equal(var.lookup(torch.tensor([9, 5])), torch.tensor([[4.0, 5.0, 6.0], [1.0, 2.0, 3.0]]))
  import pytest as _pt
  with _pt.raises(ValueError):
    sh.assign(torch.zeros(3, 3))
  ids = de.SparseIds(torch.tensor([[0, 0]]), torch.tensor([5]), (1, 1))
  sh.verify_embedding_weights(ids)
  with _pt.raises(TypeError):
    sh.verify_embedding_weights(de.SparseIds(ids.indices, torch.tensor([5], dtype=torch.int32), (1, 1)))



def test_peer_sharded_variable_with_an_eviction_strategy_python_path():
  """the Python glue of the sharded-table-with-eviction path (scripts/gpu_sharded_evict.py runs the same flow on a GPU):
  one fake shard with an LFU strategy behind PeerShardedVariable + the owner-side exchange; apply_gradients goes through
  det_peer_xchg_apply_adagrad, lookups through det_peer_xchg_find, upserts through the owner's scored insert; the one-sided
  calls are refused."""
  import numpy as np
  import torch
  from recommenders_addons_b200 import dynamic_embedding as de
  from recommenders_addons_b200._lib import DetError
  dim, slots, cap, steps = 8, 1024, 128, 30
  cfg = de.HkvHashTableConfig(init_capacity=slots, max_capacity=slots, evict_strategy=de.HkvEvictStrategy.LFU)
  var = de.Variable(dim=dim, init_size=slots, initializer=0.05, num_slot_planes=1, name="emu-shard-evict", devices=["cpu"],
                    kv_creator=de.HkvHashTableCreator(config=cfg))
  pv = de.PeerShardedVariable(fake_shards=[var])
  with pytest.raises(DetError, match="eviction strategy"):
    pv.lookup(torch.arange(4))
  nbytes = int(pv._lib.det_peer_xchg_bytes(1, cap, dim * 4))
  raw = torch.zeros(nbytes + 256, dtype=torch.uint8)
  off = (-raw.data_ptr()) % 256
  pv.attach_exchange(cap, mailbox_ptrs=[raw.data_ptr() + off], keepalive=raw, insert="push")
  pv.upsert(torch.arange(4), torch.full((4, dim), 7.0))           # through the owner: compact -> its own scored insert
  assert bool((pv.lookup(torch.arange(4)) == 7.0).all())
  opt = de.FusedAdagrad(0.1, 0.1)
  rng = np.random.default_rng(2)
  hot = torch.from_numpy(rng.choice(1 << 40, size=20, replace=False).astype(np.int64))
  f32 = np.float32
  par = {int(k): np.full(dim, 0.05, f32) for k in hot}
  acc = {int(k): np.full(dim, 0.1, f32) for k in hot}
  for t in range(steps):
    cold = torch.from_numpy((rng.choice(1 << 40, size=cap - 20, replace=False).astype(np.int64) | (1 << 41)) + t * (1 << 42))
    g = torch.from_numpy(rng.normal(0, 1e-2, (cap, dim)).astype(f32))
    pv.apply_gradients(opt, torch.cat([hot, cold]), g)
    for k, gg in zip(hot.tolist(), g[:20].numpy()):
      acc[k] = (acc[k] + gg * gg).astype(f32)
      par[k] = (par[k] - (f32(0.1) * gg) / np.sqrt(acc[k])).astype(f32)
  rows, ex = pv.lookup(hot, return_exists=True)
  assert bool(ex.all()) and np.array_equal(rows.numpy(), np.stack([par[int(k)] for k in hot]))
  st = var.tables[0].stats()
  assert st["evict_events"] > 0 and st["error_flags"] == 0 and int(var.size()) <= int(slots * 0.95)
  pv.close()

