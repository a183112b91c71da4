"""Sparse optimizer write-back (reference: python/ops/dynamic_embedding_optimizer.py:103-204, 870-958;
python/ops/embedding_weights.py:434-444).

The reference patches a stock optimizer so that, per TrainableWrapper, it reads param + slot rows from the
tables, runs the dense rule on the [N, dim] scratch and upserts param + slots back: 4 (Adagrad) or 6 (Adam)
table passes plus dense passes.  Here ONE kernel per step does find-or-insert + update in place
(det_apply_adagrad / det_apply_adam); slots live in planes co-indexed with the value rows instead of in
separate `<var>/<opt>/<slot>` tables (exportable as such through `Variable.tables[i].export(plane=k)`).
"""
import numpy as np
import torch

from .. import _lib
from .table import _ptr, _stream_ptr
from .variable import TrainableWrapper, Variable


def _init_rows(params, n, device):
  """The rows a missing key starts from: the variable's initializer, evaluated like
  Variable._create_default_values_by_initializer (one row per key when it is callable)."""
  dyn = params._create_default_values_by_initializer(n, device)
  if dyn is not None and dyn.numel() == n * params.dim:
    return dyn.reshape(n, params.dim).contiguous(), 1
  return params.tables[0]._default_value.to(device).contiguous(), 0


class _FusedBase(object):
  n_slots = 0

  def __init__(self):
    self.iterations = 0

  def _check(self, params):
    if not isinstance(params, Variable):
      raise TypeError("params should be a Variable instance.")
    if params.value_dtype != torch.float32:
      raise TypeError("fused optimizers need float32 variables")
    for t in params.tables:
      if t._num_slot_planes < self.n_slots:
        raise ValueError("variable %s was created with num_slot_planes=%d; %s needs %d" %
                         (params.name, t._num_slot_planes, type(self).__name__, self.n_slots))

  def apply_gradients(self, grads_and_vars):
    """grads_and_vars: iterable of (grad [n, dim], TrainableWrapper) or (grad, (Variable, unique ids))."""
    self.iterations += 1
    for grad, var in grads_and_vars:
      if isinstance(var, TrainableWrapper):
        params, ids = var.params, var.ids
      else:
        params, ids = var
      self._check(params)
      self.apply_sparse(params, ids.reshape(-1), grad.reshape(-1, params.dim))
      # TrainableWrapper.update_op (embedding_weights.py:441-442): the restrict policy sees every updated id
      policy = getattr(params, "restrict_policy", None)
      if policy is not None:
        policy.apply_update(ids.reshape(-1))

  def apply_sparse(self, params, keys, grads):
    """keys must be unique (they are: embedding_lookup_unique / _sparse dedupe before the lookup)."""
    grads = grads.to(torch.float32).contiguous()
    grouped, perm, bounds = params._partition(keys)
    if perm is not None:
      from .variable import gather_rows
      grads = gather_rows(grads, perm)
    for idx, (b, e) in enumerate(bounds):
      if e > b:
        self._apply_table(params, params.tables[idx], grouped[b:e].contiguous(), grads[b:e].contiguous())


class FusedAdagrad(_FusedBase):
  """TF Adagrad on dynamic-embedding rows: accum += g*g; var -= lr*g/(sqrt(accum)+epsilon).
  epsilon=0 is tf.compat.v1.train.AdagradOptimizer, epsilon=1e-7 the Keras optimizer."""
  n_slots = 1

  def __init__(self, learning_rate=0.001, initial_accumulator_value=0.1, epsilon=0.0):
    super().__init__()
    self.learning_rate = float(learning_rate)
    self.initial_accumulator_value = float(initial_accumulator_value)
    self.epsilon = float(epsilon)

  def _apply_table(self, params, table, keys, grads):
    n = keys.numel()
    init, full = _init_rows(params, n, table.device)
    _lib.check(_lib.lib().det_apply_adagrad(table.handle, _ptr(keys), _ptr(grads), n, self.learning_rate,
                                            self.epsilon, _ptr(init), full, self.initial_accumulator_value,
                                            _stream_ptr(table.device)))


class FusedAdam(_FusedBase):
  """TF Adam: m += (g-m)(1-b1); v += (g*g-v)(1-b2); var -= m*alpha/(sqrt(v)+eps) with
  alpha = lr*sqrt(1-b2^t)/(1-b1^t) (fp32, like ApplyAdam)."""
  n_slots = 2

  def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-8):
    super().__init__()
    self.learning_rate = float(learning_rate)
    self.beta_1 = float(beta_1)
    self.beta_2 = float(beta_2)
    self.epsilon = float(epsilon)

  def alpha(self):
    f = np.float32
    b1p = f(np.power(f(self.beta_1), f(self.iterations)))
    b2p = f(np.power(f(self.beta_2), f(self.iterations)))
    return float(f(f(self.learning_rate) * np.sqrt(f(1) - b2p) / (f(1) - b1p)))

  def _apply_table(self, params, table, keys, grads):
    n = keys.numel()
    init, full = _init_rows(params, n, table.device)
    _lib.check(_lib.lib().det_apply_adam(table.handle, _ptr(keys), _ptr(grads), n, self.alpha(), self.beta_1,
                                         self.beta_2, self.epsilon, _ptr(init), full, _stream_ptr(table.device)))


def DynamicEmbeddingOptimizer(self, bp_v2=False, synchronous=False, **kwargs):
  """de.DynamicEmbeddingOptimizer(optimizer): make an optimizer able to train dynamic embeddings
  (dynamic_embedding_optimizer.py:103).  Accepts a FusedAdagrad/FusedAdam (returned unchanged) or a
  torch.optim.Adagrad / Adam instance, whose hyper-parameters are carried over to the fused kernels."""
  if isinstance(self, _FusedBase):
    return self
  if isinstance(self, torch.optim.Adagrad):
    g = self.param_groups[0]
    return FusedAdagrad(g["lr"], g.get("initial_accumulator_value", 0.0), g.get("eps", 1e-10))
  if isinstance(self, torch.optim.Adam):
    g = self.param_groups[0]
    return FusedAdam(g["lr"], g["betas"][0], g["betas"][1], g["eps"])
  raise TypeError("DynamicEmbeddingOptimizer supports Adagrad and Adam on the hot path (SURVEY.md 2, #13)")
