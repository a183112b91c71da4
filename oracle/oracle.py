"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's `tfra.dynamic_embedding` hot path, used ONLY as the
parity checker by tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline /
`--impl reference` legs.  The product package (recommenders_addons_b200) never imports
this module.

Two engines with one interface:
  * RefTable  -- the reference's OWN vendored libcuckoo (+TFRA insert_or_accum) compiled
                 from /root/reference by oracle/Makefile into oracle/_ref/ (kind "reference").
  * PortTable -- oracle/cuckoo_port.c, a plain-C single-threaded restatement (kind "port").
Parity pinning: tests/test_oracle.py checks both against each other (incl. export order)
and against the reference's known-answer tests
(dynamic_embedding_variable_test.py:394-563, see tests/golden/).

NumPy restatements (all paths relative to
/root/reference/tensorflow_recommenders_addons/dynamic_embedding/):
  * default_partition_fn        python/ops/dynamic_embedding_variable.py:165-197
  * unique_first_occurrence     tf.unique as used at python/ops/dynamic_embedding_ops.py:224
  * variable_accum_values       python/ops/dynamic_embedding_variable.py:826-833
  * embedding_lookup_sparse     python/ops/dynamic_embedding_ops.py:219-291, summation order of
                                the reference test oracle kernel_tests/dynamic_embedding_ops_test.py:150-184
  * adagrad / adam dense rules  stock TF kernels invoked at python/ops/dynamic_embedding_optimizer.py:161-204
                                (TF is not vendored in the reference tree: "absolute optimizer
                                parity is unpinned"; rules restated from TF's
                                documented ApplyAdagrad(V2) / ApplyAdam update equations)
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libcuckoo_port.so")
REF_SO = os.path.join(_HERE, "_ref", "libtfra_cuckoo_ref.so")

_c_i64p = ctypes.POINTER(ctypes.c_longlong)
_c_f32p = ctypes.POINTER(ctypes.c_float)
_c_u8p = ctypes.POINTER(ctypes.c_ubyte)


def build(quiet=True):
  """Compile the checkers (port always; _ref only when /root/reference is present)."""
  out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
  if out.returncode != 0:
    raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
  if not quiet:
    print(out.stdout)


def have_ref():
  return os.path.exists(REF_SO)


def have_port():
  return os.path.exists(PORT_SO)


def _ptr(a, ty):
  return a.ctypes.data_as(ty) if a is not None else None


class _CTable:
  """Common ctypes front end; int64 keys, float32 rows."""
  _prefix = None
  _lib = None

  def __init__(self, dim, init_size=0):
    self.dim = int(dim)
    self._h = None

  # -- op semantics of cuckoo_hashtable_op.cc:211-308 ------------------------------------
  def find(self, keys, default, return_exists=False, out=None):
    """`out` (optional, [n, dim] float32) is reused as the output buffer, like TF's pooled allocator."""
    keys = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
    n = keys.shape[0]
    default = np.ascontiguousarray(default, dtype=np.float32)
    # is_full_default = (value_flat.size() == default_flat.size()), cuckoo_hashtable_op.cc:48-50
    full = int(default.size == n * self.dim)
    if not full:
      assert default.size >= self.dim, "default must hold at least one row"
    if out is None:
      out = np.empty((n, self.dim), dtype=np.float32)
    assert out.dtype == np.float32 and out.size == n * self.dim and out.flags.c_contiguous
    exists = np.zeros(n, dtype=np.uint8)
    self._fn("find")(self._h, _ptr(keys, _c_i64p), n, _ptr(default, _c_f32p), full,
                     _ptr(out, _c_f32p), _ptr(exists, _c_u8p))
    if return_exists:
      return out, exists.astype(bool)
    return out

  def insert(self, keys, values):
    keys = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
    values = np.ascontiguousarray(values, dtype=np.float32).reshape(keys.shape[0], self.dim)
    self._fn("insert")(self._h, _ptr(keys, _c_i64p), _ptr(values, _c_f32p), keys.shape[0])

  def accum(self, keys, values_or_deltas, exists):
    keys = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
    vod = np.ascontiguousarray(values_or_deltas, dtype=np.float32).reshape(keys.shape[0], self.dim)
    ex = np.ascontiguousarray(np.asarray(exists).astype(np.uint8)).reshape(-1)
    self._fn("accum")(self._h, _ptr(keys, _c_i64p), _ptr(vod, _c_f32p), _ptr(ex, _c_u8p),
                      keys.shape[0])

  def remove(self, keys):
    keys = np.ascontiguousarray(keys, dtype=np.int64).reshape(-1)
    self._fn("remove")(self._h, _ptr(keys, _c_i64p), keys.shape[0])

  def clear(self):
    self._fn("clear")(self._h)

  def size(self):
    return int(self._fn("size")(self._h))

  def export(self):
    n = self.size()
    keys = np.empty(n, dtype=np.int64)
    vals = np.empty((n, self.dim), dtype=np.float32)
    got = self._fn("export")(self._h, _ptr(keys, _c_i64p), _ptr(vals, _c_f32p), 0, n)
    return keys[:got], vals[:got]

  def import_(self, keys, values):
    """ImportValues = clear + insert (cuckoo_hashtable_op.cc:288-291)."""
    self.clear()
    self.insert(keys, values)

  def close(self):
    if self._h is not None:
      self._fn("destroy")(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def _fn(self, name):
    return getattr(type(self)._lib, self._prefix + "_" + name)


def _declare(lib, prefix, create_args):
  vp = ctypes.c_void_p
  getattr(lib, prefix + "_create").restype = vp
  getattr(lib, prefix + "_create").argtypes = create_args
  getattr(lib, prefix + "_destroy").argtypes = [vp]
  getattr(lib, prefix + "_destroy").restype = None
  getattr(lib, prefix + "_find").argtypes = [vp, _c_i64p, ctypes.c_longlong, _c_f32p, ctypes.c_int,
                                             _c_f32p, _c_u8p]
  getattr(lib, prefix + "_find").restype = None
  getattr(lib, prefix + "_insert").argtypes = [vp, _c_i64p, _c_f32p, ctypes.c_longlong]
  getattr(lib, prefix + "_insert").restype = None
  getattr(lib, prefix + "_accum").argtypes = [vp, _c_i64p, _c_f32p, _c_u8p, ctypes.c_longlong]
  getattr(lib, prefix + "_accum").restype = None
  getattr(lib, prefix + "_remove").argtypes = [vp, _c_i64p, ctypes.c_longlong]
  getattr(lib, prefix + "_remove").restype = None
  getattr(lib, prefix + "_clear").argtypes = [vp]
  getattr(lib, prefix + "_clear").restype = None
  getattr(lib, prefix + "_size").argtypes = [vp]
  getattr(lib, prefix + "_size").restype = ctypes.c_size_t
  getattr(lib, prefix + "_export").argtypes = [vp, _c_i64p, _c_f32p, ctypes.c_size_t, ctypes.c_size_t]
  getattr(lib, prefix + "_export").restype = ctypes.c_size_t


class PortTable(_CTable):
  """oracle/cuckoo_port.c (single-threaded plain-C restatement)."""
  _prefix = "port"
  kind = "port"

  def __init__(self, dim, init_size=0):
    super().__init__(dim, init_size)
    if PortTable._lib is None:
      if not have_port():
        build()
      lib = ctypes.CDLL(PORT_SO)
      _declare(lib, "port", [ctypes.c_longlong, ctypes.c_size_t])
      PortTable._lib = lib
    self.threads = 1
    self._h = PortTable._lib.port_create(self.dim, int(init_size))


class RefTable(_CTable):
  """The reference's own libcuckoo behind oracle/ref_wrapper.cc (multi-threaded like TF Shard())."""
  _prefix = "ref"
  kind = "reference"

  def __init__(self, dim, init_size=0, threads=1):
    super().__init__(dim, init_size)
    if RefTable._lib is None:
      if not have_ref():
        raise RuntimeError("oracle/_ref is not built (needs /root/reference; run `make -C oracle ref`)")
      lib = ctypes.CDLL(REF_SO)
      _declare(lib, "ref", [ctypes.c_longlong, ctypes.c_size_t, ctypes.c_int])
      lib.ref_hardware_threads.restype = ctypes.c_int
      RefTable._lib = lib
    self.threads = int(threads)
    self._h = RefTable._lib.ref_create(self.dim, int(init_size), self.threads)

  @staticmethod
  def hardware_threads():
    if RefTable._lib is None:
      RefTable(1).close()
    return int(RefTable._lib.ref_hardware_threads())


def best_table(dim, init_size=0, threads=1):
  """Reference engine when it was built here, else the port."""
  if have_ref():
    return RefTable(dim, init_size, threads)
  return PortTable(dim, init_size)


# --------------------------------------------------------------------------------------------
# NumPy restatements of the Python layer of the hot path
# --------------------------------------------------------------------------------------------
def default_partition_fn(keys, shard_num, gpu_mode=True):
  """dynamic_embedding_variable.py:165-197: (key & 0x7fffffff) % S on GPU builds, key % S else."""
  keys = np.asarray(keys, dtype=np.int64)
  if shard_num <= 1:
    return np.zeros(keys.shape, dtype=np.int32)
  if gpu_mode:
    k32 = (keys & np.int64(0x7fffffff)).astype(np.int32)
    return np.mod(k32, np.int32(shard_num)).astype(np.int32)
  return np.mod(keys, np.int64(shard_num)).astype(np.int32)  # floor-mod like tf.math.mod


def unique_first_occurrence(ids):
  """tf.unique: unique values in order of first occurrence + index of each id into them."""
  ids = np.asarray(ids).reshape(-1)
  uniq, first, inv = np.unique(ids, return_index=True, return_inverse=True)
  order = np.argsort(first, kind="stable")
  rank = np.empty_like(order)
  rank[order] = np.arange(order.shape[0])
  return uniq[order], rank[inv].astype(np.int32)


def segment_reduce(rows, idx, n_groups):
  """Gradient dedupe: TF's _deduplicate_indexed_slices = unsorted_segment_sum(values, idx, n_unique) before
  _resource_apply_sparse_duplicate_indices (dynamic_embedding_optimizer.py:150,184; also the gradient of
  dynamic_stitch / sparse_segment_sum, data_flow_grad.py:65, math_grad.py:30).  The CPU kernel adds the rows of one
  output row in increasing position, one fp32 add at a time; negative / out-of-range ids are dropped."""
  rows = np.asarray(rows, dtype=np.float32)
  idx = np.asarray(idx).reshape(-1).astype(np.int64)
  out = np.zeros((n_groups, rows.shape[1]), dtype=np.float32)
  keep = (idx >= 0) & (idx < n_groups)
  np.add.at(out, idx[keep], rows[keep])   # unbuffered, in position order
  return out


def variable_accum_values(old_values, new_values, exists):
  """Variable.accum: values_or_deltas = where(exists, new - old, new) (dynamic_embedding_variable.py:826-833)."""
  old_values = np.asarray(old_values, dtype=np.float32)
  new_values = np.asarray(new_values, dtype=np.float32)
  ex = np.asarray(exists).astype(bool).reshape(-1, 1)
  return np.where(ex, new_values - old_values, new_values).astype(np.float32)


def embedding_lookup_sparse(table, ids, segment_ids, weights, batch, combiner="mean", default=None):
  """de.embedding_lookup_sparse (dynamic_embedding_ops.py:219-291) over an oracle table.

  unique(ids) -> table.find (missing -> default row) -> gather*weights -> sequential per-segment
  sum in id order (the summation order of the reference test oracle `embedding_result`,
  dynamic_embedding_ops_test.py:150-184) -> / sum(w)  |  / sqrt(sum(w^2)).
  `default` is one row [dim] or one row per UNIQUE id [U, dim] (reference: initializer output).
  Rows of the dense result with no ids are 0 (segment_sum semantics).
  """
  ids = np.asarray(ids, dtype=np.int64).reshape(-1)
  seg = np.asarray(segment_ids, dtype=np.int64).reshape(-1)
  dim = table.dim
  uniq, idx = unique_first_occurrence(ids)
  if default is None:
    default = np.zeros(dim, dtype=np.float32)
  emb_u = table.find(uniq, default)
  emb = emb_u[idx]
  w = np.ones(ids.shape[0], dtype=np.float32) if weights is None else np.asarray(
      weights, dtype=np.float32).reshape(-1)
  out = np.zeros((batch, dim), dtype=np.float32)
  wsum = np.zeros(batch, dtype=np.float32)
  wsq = np.zeros(batch, dtype=np.float32)
  for i in range(ids.shape[0]):  # sequential, in id order, fp32 mul then fp32 add
    s = seg[i]
    out[s] = out[s] + emb[i] * w[i]
    wsum[s] = np.float32(wsum[s] + w[i])
    wsq[s] = np.float32(wsq[s] + np.float32(w[i] * w[i]))
  if combiner == "sum":
    return out
  touched = np.zeros(batch, dtype=bool)
  touched[seg] = True
  if combiner == "mean":
    div = wsum
  elif combiner == "sqrtn":
    div = np.sqrt(wsq).astype(np.float32)
  else:
    raise ValueError("combiner must be one of 'mean', 'sqrtn' or 'sum'")
  res = out.copy()
  res[touched] = out[touched] / div[touched, None]
  return res.astype(np.float32)


def adagrad_dense(param, accum, grad, lr, epsilon=0.0):
  """TF ApplyAdagrad (epsilon=0, TF1 AdagradOptimizer) / ApplyAdagradV2 (Keras, epsilon=1e-7):
       accum += g*g ; var -= lr * g / (sqrt(accum) [+ epsilon])        all fp32, no FMA."""
  f = np.float32
  g = np.asarray(grad, dtype=f)
  accum = (np.asarray(accum, dtype=f) + g * g).astype(f)
  denom = np.sqrt(accum).astype(f)
  if epsilon != 0.0:
    denom = (denom + f(epsilon)).astype(f)
  param = (np.asarray(param, dtype=f) - (f(lr) * g) / denom).astype(f)
  return param, accum


def adam_scalars(lr, beta1, beta2, step):
  """alpha = lr * sqrt(1 - beta2^t) / (1 - beta1^t), computed in fp32 like TF's ApplyAdam."""
  f = np.float32
  b1p = f(np.power(f(beta1), f(step)))
  b2p = f(np.power(f(beta2), f(step)))
  return f(f(lr) * np.sqrt(f(1) - b2p) / (f(1) - b1p))


def adam_dense(param, m, v, grad, alpha, beta1, beta2, epsilon):
  """TF ApplyAdam:  m += (g - m)*(1-b1); v += (g*g - v)*(1-b2); var -= (m*alpha)/(sqrt(v)+eps)."""
  f = np.float32
  g = np.asarray(grad, dtype=f)
  m = np.asarray(m, dtype=f)
  v = np.asarray(v, dtype=f)
  m = (m + (g - m) * (f(1) - f(beta1))).astype(f)
  v = (v + (g * g - v) * (f(1) - f(beta2))).astype(f)
  param = (np.asarray(param, dtype=f) - (m * f(alpha)) / (np.sqrt(v).astype(f) + f(epsilon))).astype(f)
  return param, m, v


def sparse_adagrad_step(param_table, accum_table, keys, grads, lr, init_param, init_accum,
                        epsilon=0.0):
  """One DynamicEmbeddingOptimizer step on unique keys (dynamic_embedding_optimizer.py:161-204):
     find(param) + find(slot)  ->  dense rule on the [U,dim] scratch  ->  upsert(param) + upsert(slot).
     Missing keys read the initializer rows (`init_param` [dim] or [U,dim]; `init_accum` likewise)."""
  keys = np.asarray(keys, dtype=np.int64).reshape(-1)
  p0 = param_table.find(keys, init_param)
  a0 = accum_table.find(keys, init_accum)
  p1, a1 = adagrad_dense(p0, a0, grads, lr, epsilon)
  param_table.insert(keys, p1)
  accum_table.insert(keys, a1)
  return p1, a1


def sparse_adam_step(param_table, m_table, v_table, keys, grads, alpha, beta1, beta2, epsilon,
                     init_param, init_m=None, init_v=None):
  keys = np.asarray(keys, dtype=np.int64).reshape(-1)
  dim = param_table.dim
  z = np.zeros(dim, dtype=np.float32)
  p0 = param_table.find(keys, init_param)
  m0 = m_table.find(keys, z if init_m is None else init_m)
  v0 = v_table.find(keys, z if init_v is None else init_v)
  p1, m1, v1 = adam_dense(p0, m0, v0, grads, alpha, beta1, beta2, epsilon)
  param_table.insert(keys, p1)
  m_table.insert(keys, m1)
  v_table.insert(keys, v1)
  return p1, m1, v1
