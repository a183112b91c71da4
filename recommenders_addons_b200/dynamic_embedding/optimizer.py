"""Sparse optimizer write-back (reference: python/ops/dynamic_embedding_optimizer.py:103-204, 870-958;
python/ops/embedding_weights.py:434-444).

The reference patches a stock optimizer so that, per TrainableWrapper, it reads param + slot rows from the
tables, runs the dense rule on the [N, dim] scratch and upserts param + slots back: 4 (Adagrad) or 6 (Adam)
table passes plus dense passes.  Here ONE kernel per step does find-or-insert + update in place
(det_apply_adagrad / det_apply_adam); slots live in planes co-indexed with the value rows instead of in
separate `<var>/<opt>/<slot>` tables (exportable as such through `Variable.tables[i].export(plane=k)`).
"""
import inspect

import numpy as np
import torch

from .. import _lib
from .table import _ptr, _stream_ptr
from .variable import TrainableWrapper, Variable, _segment_reduce_devices


def _init_rows(params, n, device):
  """The rows a missing key starts from: the variable's initializer, evaluated like
  Variable._create_default_values_by_initializer (one row per key when it is callable)."""
  dyn = params._create_default_values_by_initializer(n, device)
  if dyn is not None and dyn.numel() == n * params.dim:
    return dyn.reshape(n, params.dim).contiguous(), 1
  return params.tables[0]._default_value.to(device).contiguous(), 0


class SlotPlane(object):
  """One optimizer slot of a fused optimizer: plane `plane` of the variable's own tables, presented like the slot
  Variable the reference keeps in its own table `<var>/<opt>/<slot>` (dynamic_embedding_optimizer.py:870-958) -- enough
  of a Variable for checkpointing (`Variable.get_slot_variables(opt)` -> save / restore like any variable) and for
  inspection (`export`, `size`).  Restore the variable itself first: a slot row is only kept for keys of the table."""

  def __init__(self, params, plane, opt_name, slot_name):
    self.params, self.plane, self.slot_name = params, int(plane), slot_name
    # create_slots (dynamic_embedding_optimizer.py:882-885)
    self.name = ("%s/%s" % (params.name, slot_name)) if getattr(params, "short_file_name", False) else \
        "%s/%s/%s" % (params.name, opt_name, slot_name)
    self.dim, self.value_dtype, self.key_dtype, self.trainable = params.dim, torch.float32, params.key_dtype, False

  def size(self):
    return self.params.size()

  def export(self):
    ks, vs = zip(*[t.export(plane=self.plane) for t in self.params.tables])
    return torch.cat(ks), torch.cat(vs)

  def upsert(self, keys, values):
    """slot rows for keys of the variable (keys the variable does not hold are skipped)"""
    grouped, perm, bounds = self.params._partition(keys.reshape(-1))
    values = values.reshape(-1, self.dim).to(torch.float32)
    if perm is not None:
      from .variable import gather_rows
      values = gather_rows(values.contiguous(), perm)
    for idx, (b, e) in enumerate(bounds):
      if e > b:
        self.params.tables[idx].import_plane(self.plane, grouped[b:e].contiguous(), values[b:e].contiguous())

  def _file_name(self, idx, proc_size, proc_rank):
    return "%s_mht_%dof%d_rank%d_size%d" % (self.name.replace("/", "_"), idx + 1, self.params.shard_num, proc_rank, proc_size)

  def save_to_file_system(self, dirpath, proc_size=1, proc_rank=0, dirpath_env="TFRA_SAVED_KV", append_to_file=False,
                          buffer_size=4194304, name=None):
    for idx, t in enumerate(self.params.tables):
      t.save_plane_to_file_system(self.plane, dirpath, self._file_name(idx, proc_size, proc_rank), dirpath_env=dirpath_env,
                                  append_to_file=append_to_file, buffer_size=buffer_size)

  def load_from_file_system(self, dirpath, proc_size=1, proc_rank=0, dirpath_env="TFRA_SAVED_KV", buffer_size=4194304,
                            name=None):
    for idx, t in enumerate(self.params.tables):
      t.load_plane_from_file_system(self.plane, dirpath, self._file_name(idx, proc_size, proc_rank), dirpath_env=dirpath_env,
                                    buffer_size=buffer_size)


  def load_from_file_system_with_restore_function(self, dirpath, buffer_size=4194304):
    """reshard-on-load for a slot (the variable's own `load_from_file_system_with_restore_function` first): every
    saved shard file of this slot, whatever topology wrote it, its rows routed to the shards that hold the keys now"""
    import glob
    import os
    import re
    import numpy as np
    dirpath = os.environ.get("TFRA_SAVED_KV") or dirpath
    base = self.name.replace("/", "_")
    pat = re.compile(re.escape(base) + r"_mht_(\d+)of(\d+)_rank(\d+)_size(\d+)-keys$")
    files = sorted(f for f in glob.glob(os.path.join(dirpath, glob.escape(base) + "_mht_*-keys")) if pat.search(os.path.basename(f)))
    if not files:
      raise FileNotFoundError("no saved shards of slot %s under %s" % (self.name, dirpath))
    dev = self.params.tables[0].device
    step = max(1, int(buffer_size))
    for kf in files:
      if os.path.getsize(kf) == 0:
        continue
      keys = np.memmap(kf, dtype=np.int64, mode="r")
      vals = np.memmap(kf[:-len("-keys")] + "-values", dtype=np.float32, mode="r").reshape(-1, self.dim)
      for b in range(0, keys.shape[0], step):
        self.upsert(torch.from_numpy(np.array(keys[b:b + step])).to(dev),      # np.array: a writable copy of the chunk
                    torch.from_numpy(np.array(vals[b:b + step])).to(dev))


class _FusedBase(object):
  n_slots = 0
  opt_name = "Fused"
  slot_plane_names = ()

  def slot_names(self, dim=1):
    return list(self.slot_plane_names)

  def get_slot_names(self):
    return self.slot_names()

  def slot_variables(self, params):
    """the optimizer state of `params` as SlotPlane objects (what `Variable.get_slot_variables(opt)` returns)"""
    if not isinstance(params, Variable) or any(t._num_slot_planes < self.n_slots for t in params.tables):
      return []   # a variable without slot planes cannot hold this optimizer's state
    return [SlotPlane(params, k + 1, self.opt_name, n) for k, n in enumerate(self.slot_plane_names)]

  def get_slot(self, params, name):
    return self.slot_variables(params)[list(self.slot_plane_names).index(name)]

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
        t = params.tables[idx]   # raw pointers cross the C ABI: keys and gradients must live on the shard's own GPU
        self._apply_table(params, t, grouped[b:e].to(t.device).contiguous(), grads[b:e].to(t.device).contiguous())


  def apply_sparse_duplicate_indices(self, params, ids, grads):
    """one step from row gradients whose ids may repeat: gradients of the same id are summed first
    (`_resource_apply_sparse_duplicate_indices` -> `_deduplicate_indexed_slices`, the path the reference takes for
    IndexedSlices gradients, dynamic_embedding_optimizer.py:150,184), then the fused step runs on the unique ids.
    Single-shard fp32 variables with a static initializer take ONE C call (det_apply_*_dup): unique -> position-order
    gradient sum -> fused find-or-insert step, the unique count never leaves the device (no host synchronisation)."""
    ids = ids.reshape(-1)
    grads = grads.reshape(-1, params.dim).to(torch.float32).contiguous()
    t = params.tables[0]
    import os
    if (len(params.tables) == 1 and not callable(params.initializer) and params.dim % 4 == 0 and params.dim <= 128
        and ids.numel() > 0 and t.device.type in _segment_reduce_devices()
        and os.environ.get("DET_GRAD_REDUCE", "det") == "det"):
      ids = ids.to(t.device).contiguous()
      grads = grads.to(t.device)
      n = ids.numel()
      lib = _lib.lib()
      need = int(lib.det_apply_dup_workspace_bytes(n, params.dim))
      ws = getattr(self, "_dup_ws", None)
      if ws is None or ws.numel() < need or ws.device != t.device:
        # reused across steps (stream-ordered scratch): no allocation on the hot path after the first step
        raw = torch.empty(need + need // 8 + 256, dtype=torch.uint8, device=t.device)
        off = (-raw.data_ptr()) % 256          # the C ABI wants a 256 B aligned workspace (CUDA allocations already are)
        ws = self._dup_ws = raw[off:off + need + need // 8]
      init = t._default_value.to(t.device).contiguous()
      self._apply_table_dup(t, ids, grads, n, init, ws)
      return
    from .variable import combine_rows, unique
    uniq, idx = unique(ids)
    self.apply_sparse(params, uniq, combine_rows(grads, idx, uniq.numel()))


class FusedAdagrad(_FusedBase):
  """TF Adagrad on dynamic-embedding rows: accum += g*g; var -= lr*g/(sqrt(accum)+epsilon).
  epsilon=0 is tf.compat.v1.train.AdagradOptimizer, epsilon=1e-7 the Keras optimizer."""
  n_slots = 1
  opt_name = "Adagrad"
  slot_plane_names = ("accumulator",)

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


  def _apply_table_dup(self, table, ids, grads, n, init, ws):
    _lib.check(_lib.lib().det_apply_adagrad_dup(table.handle, _ptr(ids), _ptr(grads), n, self.learning_rate, self.epsilon,
                                                _ptr(init), self.initial_accumulator_value, _ptr(ws), ws.numel(), None,
                                                _stream_ptr(table.device)))


class FusedAdam(_FusedBase):
  """TF Adam: m += (g-m)(1-b1); v += (g*g-v)(1-b2); var -= m*alpha/(sqrt(v)+eps) with
  alpha = lr*sqrt(1-b2^t)/(1-b1^t) (fp32, like ApplyAdam)."""
  n_slots = 2
  opt_name = "Adam"
  slot_plane_names = ("m", "v")

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


  def _apply_table_dup(self, table, ids, grads, n, init, ws):
    _lib.check(_lib.lib().det_apply_adam_dup(table.handle, _ptr(ids), _ptr(grads), n, self.alpha(), self.beta_1, self.beta_2,
                                             self.epsilon, _ptr(init), _ptr(ws), ws.numel(), None,
                                             _stream_ptr(table.device)))


class ComposedOptimizer(object):
  """ANY torch.optim optimizer on dynamic-embedding rows, the way the reference patches a stock optimizer
  (dynamic_embedding_optimizer.py:137-204 `apply_grad_to_update_var`, :870-958 `create_slots`): per step and per
  TrainableWrapper, find the param rows and the rows of every slot (one slot Variable per optimizer state tensor,
  named `<var>/<Optimizer>/<slot>` like the reference's slot tables), run the stock dense rule on the [U, dim] scratch,
  write param and slots back (`TrainableWrapper.update_op`, embedding_weights.py:434-444).  Table passes per step:
  (1 + slots) finds + (1 + slots) upserts, all through the validated table kernels; Adagrad / Adam have the fused
  single-kernel path (FusedAdagrad / FusedAdam), which DynamicEmbeddingOptimizer picks by default.

  A key without slot state starts from the state a fresh torch optimizer gives a new parameter (zeros; Adagrad's
  `sum` from initial_accumulator_value) -- the reference's slot initializers.  State that is not per row (`step`,
  NAdam's `mu_product`) is carried per variable by this object, like TF's `iterations`."""

  def __init__(self, optimizer, bp_v2=False):
    if not isinstance(optimizer, torch.optim.Optimizer):
      raise TypeError("ComposedOptimizer wraps a torch.optim.Optimizer instance")
    self._cls = type(optimizer)
    # hyper-parameters of the wrapped instance; a param group may carry derived entries the constructor does not take
    accepted = set(inspect.signature(self._cls.__init__).parameters)
    self._defaults = {k: v for k, v in optimizer.param_groups[0].items() if k != "params" and k in accepted}
    self.bp_v2 = bp_v2
    self.iterations = 0
    self._slots = {}       # (variable name) -> {state key: slot Variable}
    self._slot_init = None  # state key -> initial value (float), discovered once
    self._scalar_keys = []
    self._globals = {}     # (variable name) -> state entries that are not per row

  def _discover_slots(self, dim):
    """state tensors a parameter row carries: construct the optimizer on a dummy row (state created eagerly, e.g.
    Adagrad's `sum`, gives the initial value), take one zero-gradient step (lazily created state: initial 0)"""
    p = torch.nn.Parameter(torch.zeros(2, dim + 1))   # a shape no scalar state can be mistaken for
    opt = self._cls([p], **self._defaults)
    eager = {k: float(v.reshape(-1)[0]) for k, v in opt.state.get(p, {}).items()
             if torch.is_tensor(v) and v.shape == p.shape}
    p.grad = torch.zeros_like(p)
    opt.step()
    init = {}
    for k, v in opt.state[p].items():
      if torch.is_tensor(v) and v.shape == p.shape:
        init[k] = eager.get(k, 0.0)
    self._scalar_keys = sorted(k for k in opt.state[p] if k not in init)   # step counters and other global state
    return init

  def slot_names(self, dim=1):
    if self._slot_init is None:
      self._slot_init = self._discover_slots(dim)
    return sorted(self._slot_init)

  def get_slot(self, params, name):
    """the slot Variable of `params` (dynamic_embedding_optimizer.py:870-958)"""
    return self._slots[params.name][name]

  def _slots_of(self, params):
    if self._slot_init is None:
      self._slot_init = self._discover_slots(params.dim)
    if params.name not in self._slots:
      made = {}
      for k, init in self._slot_init.items():
        made[k] = Variable(key_dtype=params.key_dtype, value_dtype=params.value_dtype, dim=params.dim,
                           devices=[str(d) for d in params.devices], partitioner=params.partition_fn,
                           name=("%s/%s" % (params.name, k)) if getattr(params, "short_file_name", False) else
                           "%s/%s/%s" % (params.name, self._cls.__name__, k), initializer=init, trainable=False,
                           init_size=params.init_size, kv_creator=params.kv_creator)
      self._slots[params.name] = made
      # create_slots (dynamic_embedding_optimizer.py:870-958): a restrict policy shrinks the slot tables with the variable
      if getattr(params, "restrict_policy", None) is not None:
        params.restrict_policy._track_params_from_optimizer_slots(list(made.values()))
    return self._slots[params.name]

  def get_slot_names(self):
    return self.slot_names()

  def slot_variables(self, params):
    """every slot Variable this optimizer keeps for `params` (empty before the first step)"""
    return [v for _, v in sorted(self._slots.get(params.name, {}).items())]

  @staticmethod
  def _materialize(opt, p):
    """torch's INITIAL state of `p` without taking a step (most optimizers create it lazily inside step()): their
    `_init_group(group, params_with_grad, grads, *state lists)` does exactly that.  Best effort (private API)."""
    try:
      n_lists = len(inspect.signature(opt._init_group).parameters) - 1
      opt._init_group(opt.param_groups[0], *[[] for _ in range(n_lists)])
    except Exception:  # pylint: disable=broad-except
      pass
    return opt.state.get(p, {})

  def apply_gradients(self, grads_and_vars):
    """grads_and_vars: iterable of (grad [n, dim], TrainableWrapper) or (grad, (Variable, unique ids))."""
    self.iterations += 1
    for grad, var in grads_and_vars:
      tw = var if isinstance(var, TrainableWrapper) else TrainableWrapper(var[0], var[1])
      self._apply_one(tw, grad)

  def apply_sparse(self, params, keys, grads):
    """one step on the rows of `params` selected by the unique `keys` (the entry point the sharded variables call
    on the owning rank after routing and combining the gradients; same name as the fused optimizers')"""
    self._apply_one(TrainableWrapper(params, keys), grads)

  def apply_sparse_duplicate_indices(self, params, ids, grads):
    """as the fused optimizers': gradients of repeated ids are summed (`_deduplicate_indexed_slices`), then one step"""
    from .variable import combine_rows, unique
    uniq, idx = unique(ids.reshape(-1))
    self.apply_sparse(params, uniq, combine_rows(grads.reshape(-1, params.dim), idx, uniq.numel()))

  def _apply_one(self, tw, grad):
    params = tw.params
    if not isinstance(params, Variable):
      raise TypeError("params should be a Variable instance.")
    flat = tw.ids.reshape(-1)
    if tw.values is None or (tw._old_values is None and params.bp_v2):
      tw.prefetch_values()
    p = torch.nn.Parameter(tw.values.detach().clone().reshape(-1, params.dim))
    p.grad = grad.reshape(-1, params.dim).to(p.dtype)
    opt = self._cls([p], **self._defaults)
    slots = self._slots_of(params)
    # state that is not per row (step counters, NAdam's mu_product ...) is carried by this object, per variable
    carried = self._globals.get(params.name)
    if carried is None:
      carried = {k: v for k, v in self._materialize(opt, p).items() if k not in slots}
    inject = bool(carried) or not self._scalar_keys
    if inject:   # (else: first step of an optimizer whose initial scalars are unknown -> torch's fresh state,
      #             identical to the slot tables' initial values)
      state = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in carried.items()}
      for k, sv in slots.items():
        state[k] = sv.lookup(flat).reshape(-1, params.dim).clone()
      opt.state[p] = state
    opt.step()
    after = opt.state[p]
    self._globals[params.name] = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in after.items()
                                  if k not in slots}
    tw.values = p.detach()
    tw.update_op()
    for k, sv in slots.items():
      if torch.is_tensor(after.get(k)):
        sv.upsert(flat, after[k])


def DynamicEmbeddingOptimizer(self, bp_v2=False, synchronous=False, fused=True, **kwargs):
  """de.DynamicEmbeddingOptimizer(optimizer): make an optimizer able to train dynamic embeddings
  (dynamic_embedding_optimizer.py:103).  Accepts a FusedAdagrad/FusedAdam (returned unchanged) or any
  torch.optim.Optimizer instance: Adagrad / Adam are carried over to the fused single-kernel path (fused=True, the
  default); every other optimizer -- or fused=False -- gets the reference's composed find -> dense rule -> upsert
  path (ComposedOptimizer)."""
  if isinstance(self, (_FusedBase, ComposedOptimizer)):
    return self
  if fused and type(self) is torch.optim.Adagrad:
    g = self.param_groups[0]
    if not (g.get("lr_decay", 0) or g.get("weight_decay", 0) or g.get("maximize", False)):
      return FusedAdagrad(g["lr"], g.get("initial_accumulator_value", 0.0), g.get("eps", 1e-10))
  if fused and type(self) is torch.optim.Adam:
    g = self.param_groups[0]
    if not (g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False)):
      return FusedAdam(g["lr"], g["betas"][0], g["betas"][1], g["eps"])
  if isinstance(self, torch.optim.Optimizer):
    return ComposedOptimizer(self, bp_v2=bp_v2)
  raise TypeError("DynamicEmbeddingOptimizer needs a torch.optim.Optimizer, de.FusedAdagrad or de.FusedAdam")
