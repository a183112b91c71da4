"""Callers either side of the fused kernels, composed from the validated table ops (SURVEY 8a rows a13 / a15 / a16's
single-GPU part): `TrainableWrapper` (prefetch -> dense scratch -> update_op, with and without bp_v2), `ModelMode`,
`shadow_ops.ShadowVariable`, and `DynamicEmbeddingOptimizer` over ANY stock optimizer (ComposedOptimizer: one slot
Variable per optimizer state, find -> dense rule -> upsert) -- tested the way the reference tests its optimizer patch:
a twin model on a plain dense parameter trained with the unpatched optimizer, 10 steps, compared at fp32 tolerance
(kernel_tests/dynamic_embedding_optimizer_test.py:349-440, swept over the optimizer list of :112-278).

First hardware run: round 1's driver box (all five suites passed on a fresh B200); ungated in round 2."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"   # tests/test_mirror_emu.py re-runs these bodies over the emulated library with DEV = "cpu"
V, DIM, STEPS = 48, 6, 10

OPTIMIZERS = {
    "sgd": lambda p: torch.optim.SGD(p, lr=0.1),
    "momentum": lambda p: torch.optim.SGD(p, lr=0.1, momentum=0.9),
    "nesterov": lambda p: torch.optim.SGD(p, lr=0.1, momentum=0.9, nesterov=True),
    "rmsprop": lambda p: torch.optim.RMSprop(p, lr=0.01),
    "rmsprop_centered_momentum": lambda p: torch.optim.RMSprop(p, lr=0.01, centered=True, momentum=0.5),
    "adagrad": lambda p: torch.optim.Adagrad(p, lr=0.1, initial_accumulator_value=0.1),
    "adadelta": lambda p: torch.optim.Adadelta(p, lr=1.0),
    "adam": lambda p: torch.optim.Adam(p, lr=0.01),
    "adam_amsgrad": lambda p: torch.optim.Adam(p, lr=0.01, amsgrad=True),
    "adamw": lambda p: torch.optim.AdamW(p, lr=0.01, weight_decay=0.01),
    "adamax": lambda p: torch.optim.Adamax(p, lr=0.01),
    "nadam": lambda p: torch.optim.NAdam(p, lr=0.01),
}


def _de():
  from recommenders_addons_b200 import dynamic_embedding as de
  return de


def _ids_stream(rng):
  """every step looks ALL rows up (some twice): a dense optimizer then touches exactly the rows the sparse one does"""
  out = []
  for _ in range(STEPS):
    ids = np.concatenate([rng.permutation(V), rng.integers(0, V, 16)]).astype(np.int64)
    out.append(ids.reshape(4, -1))
  return out


def _train_twin(make_opt, stream, weights, init):
  dense = torch.nn.Parameter(torch.full((V, DIM), init, device=DEV))
  opt = make_opt([dense])
  for ids, w in zip(stream, weights):
    opt.zero_grad()
    emb = dense[torch.as_tensor(ids, device=DEV)]
    (emb * w).sum().backward()
    opt.step()
  return dense.detach()


@pytest.mark.parametrize("name", ["momentum", "rmsprop", "adam", "adagrad"])
def test_composed_optimizer_bp_v2_matches_dense_twin(name):
  """the reference sweeps every optimizer with bp_v2 on and off (dynamic_embedding_optimizer_test.py:112-278): with
  bp_v2 the write-back is accum(old, new, exists) -- the same result when nobody else writes the rows meanwhile"""
  test_composed_optimizer_matches_dense_twin(name, bp_v2=True)


@pytest.mark.parametrize("name", sorted(OPTIMIZERS))
def test_composed_optimizer_matches_dense_twin(name, bp_v2=False):
  de = _de()
  rng = np.random.default_rng(3)
  stream = _ids_stream(rng)
  weights = [torch.as_tensor(rng.normal(0, 1, s.shape + (DIM,)).astype(np.float32), device=DEV) for s in stream]
  twin = _train_twin(OPTIMIZERS[name], stream, weights, 0.25)
  var = de.get_variable("twin-%s-%d" % (name, bp_v2), dim=DIM, initializer=0.25, devices=[DEV], bp_v2=bp_v2)
  opt = de.DynamicEmbeddingOptimizer(OPTIMIZERS[name]([torch.nn.Parameter(torch.zeros(1))]), fused=False, bp_v2=bp_v2)
  assert isinstance(opt, de.ComposedOptimizer)
  for ids, w in zip(stream, weights):
    emb, tw = de.embedding_lookup_unique(var, torch.as_tensor(ids, device=DEV), return_trainable=True)
    (emb * w).sum().backward()
    opt.apply_gradients([(tw.values.grad, tw)])
  got = var.lookup(torch.arange(V, device=DEV))
  assert int(var.size()) == V
  # bp_v2 adds (new - old) to the stored row: one more rounding per step than the plain write-back
  tol = 1e-5 if bp_v2 else 2e-6
  np.testing.assert_allclose(got.cpu().numpy(), twin.cpu().numpy(), rtol=tol, atol=tol)
  for slot in opt.slot_names():
    assert int(opt.get_slot(var, slot).size()) == V      # one slot table per optimizer state, like create_slots


@pytest.mark.parametrize("kind", ["adagrad", "adam"])
def test_composed_and_fused_paths_agree(kind):
  """the fused single-kernel step (TF rule) and the composed path over torch's rule, same stream: 1e-6 -- Adagrad is
  the same rule; for Adam torch applies epsilon after the bias correction (tests/test_oracle.py), hence eps = 0"""
  de = _de()
  rng = np.random.default_rng(5)
  stream = _ids_stream(rng)
  weights = [torch.as_tensor(rng.normal(0, 1, s.shape + (DIM,)).astype(np.float32), device=DEV) for s in stream]
  if kind == "adagrad":
    fused, stock = de.FusedAdagrad(0.1, 0.1, 1e-10), torch.optim.Adagrad([torch.nn.Parameter(torch.zeros(1))], lr=0.1,
                                                                         initial_accumulator_value=0.1)
  else:
    fused, stock = de.FusedAdam(0.01, 0.9, 0.999, 0.0), torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=0.01,
                                                                         eps=0.0)
  assert type(de.DynamicEmbeddingOptimizer(stock)).__name__ == type(fused).__name__   # the default: the fused path
  res = []
  for tag, opt in (("f", fused), ("c", de.DynamicEmbeddingOptimizer(stock, fused=False))):
    var = de.get_variable("agree-%s-%s" % (kind, tag), dim=DIM, initializer=0.25, devices=[DEV], num_slot_planes=2)
    for ids, w in zip(stream, weights):
      emb, tw = de.embedding_lookup_unique(var, torch.as_tensor(ids, device=DEV), return_trainable=True)
      (emb * w).sum().backward()
      opt.apply_gradients([(tw.values.grad, tw)])
    res.append(var.lookup(torch.arange(V, device=DEV)).cpu().numpy())
  np.testing.assert_allclose(res[0], res[1], rtol=1e-5, atol=1e-6)


def test_untouched_rows_and_slot_tables_stay_sparse():
  de = _de()
  var = de.get_variable("sparse-touch", dim=DIM, initializer=1.0, devices=[DEV])
  var.upsert(torch.arange(100, device=DEV), torch.ones(100, DIM, device=DEV))
  opt = de.DynamicEmbeddingOptimizer(torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.5, momentum=0.9))
  ids = torch.tensor([3, 7, 7, 200], device=DEV)
  for _ in range(2):
    emb, tw = de.embedding_lookup_unique(var, ids, return_trainable=True)
    emb.sum().backward()
    opt.apply_gradients([(tw.values.grad, tw)])
  got = var.lookup(torch.tensor([3, 7, 200, 5], device=DEV)).cpu()
  # momentum SGD, lr .5: grads 1 / 2 / 1 per step (id 7 appears twice); buf1 = g, buf2 = .9 g + g
  exp = lambda g: 1.0 - 0.5 * g - 0.5 * (0.9 * g + g)   # noqa: E731
  assert torch.allclose(got, torch.tensor([[exp(1.0)] * DIM, [exp(2.0)] * DIM, [exp(1.0)] * DIM, [1.0] * DIM]))
  assert int(var.size()) == 101                                        # id 200 was created by its first update
  assert int(opt.get_slot(var, "momentum_buffer").size()) == 3         # only the touched keys carry slot state


def test_get_slot_variables_and_restrict_policy_tracks_them():
  """kernel_tests/dynamic_embedding_variable_test.py:1959-2002 (get_slot_variables) + create_slots' hand-over of the
  slot variables to the restrict policy (dynamic_embedding_optimizer.py:870-958): restricting the variable shrinks its
  slot tables too"""
  de = _de()
  var = de.get_variable("trl4397", dim=DIM, initializer=0.0, devices=[DEV], restrict_policy=de.TimestampRestrictPolicy)
  opt = de.DynamicEmbeddingOptimizer(torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=0.1), fused=False)
  assert var.get_slot_variables(opt) == [] and var.get_slot_variables(de.FusedAdam(0.1)) == []
  with pytest.raises(TypeError):
    var.get_slot_variables(object())
  ids = torch.arange(1, 9, device=DEV)
  emb, tw = de.embedding_lookup(var, ids, return_trainable=True)
  emb.sum().backward()
  opt.apply_gradients([(tw.values.grad, tw)])
  names = [v.name for v in var.get_slot_variables(opt)]
  assert names == ["trl4397/Adam/exp_avg", "trl4397/Adam/exp_avg_sq"] and opt.get_slot_names() == ["exp_avg", "exp_avg_sq"]
  assert [int(v.size()) for v in var.get_slot_variables(opt)] == [8, 8] and int(var.restrict_policy.status.size()) == 8
  assert [id(p) for p in var.restrict_policy.params_in_slots] == [id(v) for v in var.get_slot_variables(opt)]
  var.restrict(3)
  assert int(var.size()) == 3 and [int(v.size()) for v in var.get_slot_variables(opt)] == [3, 3]
  sh = de.shadow_ops.ShadowVariable(var, name="user_embedding")
  assert var.get_trainable_by_name("user_embedding") is sh and var.get_trainable_by_name("nope") is None


def test_layers_and_apply_sparse_take_any_optimizer():
  """de.layers.Embedding.apply_gradients with a composed optimizer; ComposedOptimizer.apply_sparse is what the sharded
  variables call on the owning rank (same name and arguments as the fused optimizers')"""
  de = _de()
  layer = de.layers.Embedding(DIM, initializer=1.0, devices=[DEV], name="layer-any-opt", num_slot_planes=0)
  opt = de.DynamicEmbeddingOptimizer(torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=0.5))
  layer.train()
  layer(torch.tensor([[3, 4], [4, 5]], device=DEV)).sum().backward()
  layer.apply_gradients(opt)
  got = layer.params.lookup(torch.tensor([3, 4, 5, 6], device=DEV)).cpu()
  assert torch.allclose(got[:, 0], torch.tensor([0.5, 0.0, 0.5, 1.0]))    # id 4 appears twice: gradient 2
  opt.apply_sparse(layer.params, torch.tensor([6, 7], device=DEV), torch.ones(2, DIM, device=DEV))
  assert torch.allclose(layer.params.lookup(torch.tensor([6, 7], device=DEV)).cpu(), torch.full((2, DIM), 0.5))


def test_model_mode_and_trainable_wrapper_filter():
  de = _de()
  var = de.get_variable("modes", dim=DIM, initializer=2.0, devices=[DEV])
  ids = torch.tensor([[1, 2], [2, 9]], device=DEV)
  assert de.get_model_mode() == de.ModelMode.TRAIN
  try:
    de.enable_inference_mode()
    assert de.get_model_mode() == "inference"
    emb, tw = de.embedding_lookup(var, ids, return_trainable=True)
    assert emb.shape == (2, 2, DIM) and not emb.requires_grad and not tw.values.requires_grad
    tw.values = tw.values + 1.0
    tw.update_op()                                   # read-only in inference mode
    assert int(var.size()) == 0
    sh = de.shadow_ops.ShadowVariable(var, name="modes-shadow")
    out = de.shadow_ops.embedding_lookup(sh, ids)
    assert out.shape == (2, 2, DIM) and not out.requires_grad and sh.ids.numel() == 0
  finally:
    de.enable_train_mode()
  # test_inference_numberic_correctness (dynamic_embedding_optimizer_test.py:1506-1562): the same prediction in both
  # modes; a one-element list of params is accepted, more is an error (dynamic_embedding_variable.py:1403-1406)
  var.upsert(torch.arange(20, device=DEV), torch.rand(20, DIM, device=DEV))
  preds = []
  for fn in (de.enable_train_mode, de.enable_inference_mode):
    fn()
    test_var, _ = de.embedding_lookup([var], torch.tensor([0, 1, 2, 3, 4], device=DEV), return_trainable=True)
    preds.append((test_var.detach() + 1) * 1.0)
  de.enable_train_mode()
  assert torch.equal(preds[0], preds[1])
  with pytest.raises(ValueError):
    de.embedding_lookup([var, var], ids)
  emb, tw = de.embedding_lookup(var, ids, return_trainable=True)
  assert tw.values.requires_grad and tw.model_mode == "train"
  assert de.trainable_wrapper_filter([tw, var, 3, sh]) == [tw, sh]


@pytest.mark.parametrize("bp_v2", [False, True])
def test_trainable_wrapper_prefetch_and_update_op(bp_v2):
  """embedding_weights.py:163-170 / :434-444: prefetch -> step the scratch -> update_op; with bp_v2 the write-back is
  accum(old, new, exists): new keys are inserted with their value, resident keys get the DIFFERENCE added"""
  de = _de()
  var = de.get_variable("tw-%d" % bp_v2, dim=DIM, initializer=0.5, devices=[DEV], bp_v2=bp_v2)
  var.upsert(torch.tensor([1, 2], device=DEV), torch.tensor([[1.0] * DIM, [2.0] * DIM], device=DEV))
  tw = de.TrainableWrapper(var, torch.tensor([2, 1, 40], device=DEV))
  vals = tw.prefetch_values()
  assert torch.equal(vals.cpu(), torch.tensor([[2.0] * DIM, [1.0] * DIM, [0.5] * DIM]))
  if bp_v2:
    assert tw.exists.cpu().tolist() == [True, True, False]
    var.upsert(torch.tensor([1], device=DEV), torch.full((1, DIM), 10.0, device=DEV))   # a concurrent writer
  vals.sum().backward()
  with torch.no_grad():
    tw.values -= 0.25 * tw.values.grad
  tw.update_op()
  got = var.lookup(torch.tensor([1, 2, 40], device=DEV)).cpu()
  row1 = 10.0 - 0.25 if bp_v2 else 0.75            # bp_v2 adds the delta to whatever is in the table NOW
  assert torch.allclose(got, torch.tensor([[row1] * DIM, [1.75] * DIM, [0.25] * DIM]))
  assert int(tw.size()) == 3


def test_shadow_variable_training_matches_dense_twin():
  """shadow_embedding_ops.py:242-350 + keras/layers/embedding.py:287-330: the layer's path -- shadow lookup (unique),
  loss, gradient of the shadow's scratch, optimizer, write-back"""
  de = _de()
  rng = np.random.default_rng(9)
  stream = _ids_stream(rng)
  weights = [torch.as_tensor(rng.normal(0, 1, s.shape + (DIM,)).astype(np.float32), device=DEV) for s in stream]
  make = OPTIMIZERS["rmsprop"]
  twin = _train_twin(make, stream, weights, 0.25)
  var = de.get_variable("shadow-twin", dim=DIM, initializer=0.25, devices=[DEV])
  shadow = de.shadow_ops.ShadowVariable(var, name="shadow-twin/shadow")
  assert var._trainable_store["shadow-twin/shadow"] is shadow
  opt = de.DynamicEmbeddingOptimizer(make([torch.nn.Parameter(torch.zeros(1))]))
  for ids, w in zip(stream, weights):
    ids = torch.as_tensor(ids, device=DEV)
    emb = de.shadow_ops.embedding_lookup_unique(shadow, ids, DIM, with_unique=True)
    assert emb.shape == tuple(ids.shape) + (DIM,) and shadow.ids.numel() == V
    (emb * w).sum().backward()
    opt.apply_gradients([(shadow.values.grad, shadow)])
  got = var.lookup(torch.arange(V, device=DEV))
  np.testing.assert_allclose(got.cpu().numpy(), twin.cpu().numpy(), rtol=2e-6, atol=2e-6)
  with pytest.raises(ValueError):
    de.shadow_ops.embedding_lookup(shadow, torch.tensor([1], dtype=torch.int32, device=DEV))
  with pytest.raises(TypeError):
    de.shadow_ops.ShadowVariable(object())


@pytest.mark.parametrize("dim", [8, 64, 128])
def test_lookup_sparse_max_norm_fused_vs_composed(dim):
  """det_lookup_sparse_clip (max_norm folded into the gather) against the composed path, 1e-6"""
  de = _de()
  rng = np.random.default_rng(dim)
  var = de.get_variable("mn-gpu-%d" % dim, dim=dim, initializer=0.0, devices=[DEV])
  keys = torch.arange(5000, device=DEV)
  rows = torch.as_tensor(rng.normal(0, 1.0 / np.sqrt(dim), (5000, dim)).astype(np.float32), device=DEV) * \
      torch.linspace(0.2, 3, 5000, device=DEV)[:, None]
  var.upsert(keys, rows)
  n, batch = 40000, 6000
  ind = torch.stack([torch.sort(torch.as_tensor(rng.integers(0, batch, n), device=DEV)).values, torch.arange(n, device=DEV)], 1)
  sp = de.SparseIds(ind, torch.as_tensor(rng.integers(0, 6000, n), device=DEV), (batch, n))
  sw = de.SparseIds(ind, torch.as_tensor(rng.uniform(0.5, 2, n).astype(np.float32), device=DEV), (batch, n))
  for comb in ("sum", "mean", "sqrtn"):
    got = de.embedding_lookup_sparse(var, sp, sw, combiner=comb, max_norm=1.0)
    composed, _ = de.embedding_lookup_sparse(var, sp, sw, combiner=comb, max_norm=1.0, return_trainable=True)
    np.testing.assert_allclose(got.cpu().numpy(), composed.detach().cpu().numpy(), rtol=2e-6, atol=2e-6)


def test_read_only_ops_are_cuda_graph_capturable():
  """serving: det_find / det_lookup_sparse take no host lock, allocate nothing and never synchronise, so a caller can
  capture them in a CUDA graph (small-batch inference is launch-bound) and replay with new keys in the same buffer"""
  if DEV != "cuda":
    pytest.skip("CUDA graphs need a GPU")
  de = _de()
  from recommenders_addons_b200.dynamic_embedding.ops import lookup_sparse_fused
  dim, n = 64, 4096
  var = de.get_variable("graph-capture", dim=dim, initializer=0.5, devices=[DEV])
  rng = np.random.default_rng(1)
  keys = np.unique(rng.integers(0, 1 << 40, 3 * n))[:2 * n].astype(np.int64)
  vals = rng.normal(0, 0.01, (len(keys), dim)).astype(np.float32)
  var.upsert(K(keys), torch.as_tensor(vals, device=DEV))
  table = var.tables[0]
  static_keys = K(keys[:n]).clone()
  seg = torch.arange(n, device=DEV, dtype=torch.int32) // 4
  side = torch.cuda.Stream()
  with torch.cuda.stream(side):          # warm-up: occupancy queries, scratch buffers
    for _ in range(3):
      table.lookup(static_keys, return_exists=True)
      lookup_sparse_fused(var, static_keys, seg, None, n // 4, "sum")
  torch.cuda.current_stream().wait_stream(side)
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    out, ex = table.lookup(static_keys, return_exists=True)
    pooled = lookup_sparse_fused(var, static_keys, seg, None, n // 4, "sum")
  for trial in range(3):
    q = np.concatenate([keys[rng.permutation(len(keys))[:n - 100]], rng.integers(1 << 41, 1 << 42, 100)]).astype(np.int64)
    static_keys.copy_(K(q))
    graph.replay()
    torch.cuda.synchronize()
    exp, eex = table.lookup(K(q), return_exists=True)
    assert torch.equal(out, exp) and torch.equal(ex, eex)
    assert torch.equal(pooled, lookup_sparse_fused(var, K(q), seg, None, n // 4, "sum"))


def K(a):
  return torch.as_tensor(np.asarray(a, dtype=np.int64), device=DEV)


@pytest.mark.parametrize("kind", ["adagrad", "adam"])
def test_fused_optimizer_state_survives_a_checkpoint(kind, tmp_path):
  """checkpoint round trip of a training run (dynamic_embedding_optimizer_test.py:1262-1470): the variable AND the
  optimizer slots are saved (`Variable.get_slot_variables(opt)`, one raw file pair per slot like the reference's slot
  tables), restored into a fresh variable, and training continues BIT-IDENTICALLY to the uninterrupted run"""
  de = _de()
  rng = np.random.default_rng(9)
  steps = [(rng.choice(V, 12, replace=False).astype(np.int64), rng.normal(0, 1, (12, DIM)).astype(np.float32)) for _ in range(6)]
  make = (lambda: de.FusedAdagrad(0.1, 0.1)) if kind == "adagrad" else (lambda: de.FusedAdam(0.01))
  planes = 1 if kind == "adagrad" else 2

  def run(var, opt, part):
    for ids, g in part:
      opt.apply_gradients([(torch.as_tensor(g, device=DEV), (var, torch.as_tensor(ids, device=DEV)))])

  ref = de.get_variable("ckpt-ref-" + kind, dim=DIM, initializer=0.25, devices=[DEV], num_slot_planes=planes)
  ropt = make()
  run(ref, ropt, steps)
  a = de.get_variable("ckpt-" + kind, dim=DIM, initializer=0.25, devices=[DEV], num_slot_planes=planes)
  aopt = make()
  run(a, aopt, steps[:3])
  slots = a.get_slot_variables(aopt)
  assert [s.slot_name for s in slots] == (["accumulator"] if kind == "adagrad" else ["m", "v"])
  assert slots[0].name == "ckpt-%s/%s/%s" % (kind, "Adagrad" if kind == "adagrad" else "Adam", slots[0].slot_name)
  a.save_to_file_system(str(tmp_path))
  for s in slots:
    s.save_to_file_system(str(tmp_path))
  # a fresh process: new variable of the same name and topology, new optimizer object with the step counter restored
  from recommenders_addons_b200.dynamic_embedding import variable as VM
  VM._VARIABLES.pop("ckpt-" + kind)
  b = de.get_variable("ckpt-" + kind, dim=DIM, initializer=0.25, devices=[DEV], num_slot_planes=planes)
  bopt = make()
  bopt.iterations = aopt.iterations
  b.load_from_file_system(str(tmp_path))
  for s in b.get_slot_variables(bopt):
    s.load_from_file_system(str(tmp_path))
  for sa, sb in zip(slots, b.get_slot_variables(bopt)):
    ka, va = sa.export()
    kb, vb = sb.export()
    oa, ob = torch.argsort(ka), torch.argsort(kb)
    assert torch.equal(ka[oa], kb[ob]) and torch.equal(va[oa], vb[ob])
  run(b, bopt, steps[3:])
  q = torch.arange(V, device=DEV)
  assert torch.equal(b.lookup(q), ref.lookup(q))
  # a slot row can also be written directly (restore from another source); keys the variable does not hold are skipped
  sp = b.get_slot_variables(bopt)[0]
  sp.upsert(torch.as_tensor([int(steps[0][0][0]), 10**9], device=DEV), torch.full((2, DIM), 7.0, device=DEV))
  k, v = sp.export()
  assert bool((v[k == int(steps[0][0][0])] == 7.0).all()) and int(b.size()) == int(k.numel())


def test_file_system_saver_checkpoints_optimizer_state_across_a_reshard(tmp_path):
  """FileSystemSaver.save / restore with `optimizer=`: a 2-shard variable and its Adam slots are written shard by
  shard, restored into a 3-shard variable (reshard-on-load for the rows AND the slots), and training continues
  bit-identically to an uninterrupted single-shard run"""
  de = _de()
  rng = np.random.default_rng(19)
  steps = [(rng.choice(V, 14, replace=False).astype(np.int64), rng.normal(0, 1, (14, DIM)).astype(np.float32)) for _ in range(6)]

  def run(var, opt, part):
    for ids, g in part:
      opt.apply_gradients([(torch.as_tensor(g, device=DEV), (var, torch.as_tensor(ids, device=DEV)))])

  ref = de.get_variable("fss-ref", dim=DIM, initializer=0.5, devices=[DEV], num_slot_planes=2)
  ropt = de.FusedAdam(0.01)
  run(ref, ropt, steps)
  a = de.get_variable("fss", dim=DIM, initializer=0.5, devices=[DEV] * 2, num_slot_planes=2)
  aopt = de.FusedAdam(0.01)
  run(a, aopt, steps[:3])
  saver = de.FileSystemSaver(save_path=str(tmp_path))
  saver.save(a, optimizer=aopt)
  import os
  names = sorted(os.listdir(str(tmp_path)))
  assert "fss_mht_1of2_rank0_size1-keys" in names and "fss_Adam_m_mht_2of2_rank0_size1-values" in names and len(names) == 12
  from recommenders_addons_b200.dynamic_embedding import variable as VM
  VM._VARIABLES.pop("fss")
  b = de.get_variable("fss", dim=DIM, initializer=0.5, devices=[DEV] * 3, num_slot_planes=2)
  bopt = de.FusedAdam(0.01)
  bopt.iterations = aopt.iterations
  saver.restore(b, optimizer=bopt)
  assert int(b.size()) == int(a.size())
  run(b, bopt, steps[3:])
  q = torch.arange(V, device=DEV)
  assert torch.equal(b.lookup(q), ref.lookup(q))


def test_restrict_shrinks_the_fused_optimizer_slots_with_the_variable():
  """restrict_policies_test.py (`*_apply_restriction`: after the restriction the variable AND its slot tables hold the
  reserved keys only): with the fused optimizers the slots are planes of the variable's own table, so removing a key
  removes its optimizer state with it -- and a key that comes back starts from the slot initializer again"""
  de = _de()
  var = de.get_variable("restrict-fused", dim=DIM, initializer=0.0, devices=[DEV], num_slot_planes=1,
                        restrict_policy=de.FrequencyRestrictPolicy)
  opt = de.FusedAdagrad(0.1, 0.1)
  ids = torch.arange(0, 12, device=DEV)
  for rep in range(3):      # ids 0..3 are seen three times, 4..7 twice, 8..11 once
    sel = ids[: 12 - 4 * rep]
    opt.apply_gradients([(torch.ones(sel.numel(), DIM, device=DEV), (var, sel))])
  slot = var.get_slot_variables(opt)[0]
  assert int(var.size()) == 12 and int(slot.export()[0].numel()) == 12
  var.restrict(4, trigger=4)
  k, a = slot.export()
  assert sorted(k.tolist()) == [0, 1, 2, 3] and int(var.size()) == 4
  assert torch.allclose(a, torch.full_like(a, 0.1 + 3.0))                      # three steps of g = 1: a = 0.1 + 3 * 1
  opt.apply_gradients([(torch.ones(1, DIM, device=DEV), (var, torch.tensor([9], device=DEV)))])   # key 9 returns
  k, a = slot.export()
  assert torch.allclose(a[k == 9], torch.full((1, DIM), 0.1 + 1.0, device=a.device))
