"""`de.embedding_lookup_sparse` / `de.safe_embedding_lookup_sparse`
(reference: python/ops/dynamic_embedding_ops.py:120-438)."""
import torch

from .. import _lib
from .table import _ptr, _stream_ptr
from .variable import Variable, embedding_lookup, gather_unique, unique


class SparseIds(object):
  """Minimal stand-in for tf.SparseTensor: `indices` [nnz, rank] int64 in canonical row-major order,
  `values` [nnz], `dense_shape` (tuple)."""

  def __init__(self, indices, values, dense_shape):
    self.indices = indices
    self.values = values
    self.dense_shape = tuple(int(d) for d in dense_shape)


def _table_device_types():
  from . import table
  return table._DEVICE_TYPES   # ("cuda",): the library only ever sees device memory


def _fused_ok(params, default_rows):
  return (params.shard_num == 1 and params.value_dtype == torch.float32 and default_rows.numel() == params.dim)


def _clip_fusable(dim):
  """rows the fused max_norm covers (det_lookup_sparse_clip): one vector per lane of a 32-lane group"""
  return dim <= 128 if dim % 4 == 0 else dim <= 32


def lookup_sparse_fused(params, ids, segment_ids, weights, batch, combiner, default_row=None, max_norm=None):
  """K6: one kernel does find -> [clip] -> *weight -> per-segment sum -> normalise (det_lookup_sparse[_clip])."""
  table = params.tables[0]
  dev = table.device
  ids = ids.reshape(-1).to(dev).contiguous()
  seg = segment_ids.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
  w = None if weights is None else weights.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
  if default_row is None:
    default_row = table._default_value
  default_row = default_row.to(device=dev, dtype=torch.float32).contiguous()
  out = torch.empty((batch, params.dim), dtype=torch.float32, device=dev)
  if max_norm is not None:
    _lib.check(_lib.lib().det_lookup_sparse_clip(table.handle, _ptr(ids), _ptr(seg), _ptr(w), ids.numel(), batch,
                                                 _lib.COMBINERS[combiner], _ptr(default_row), float(max_norm), _ptr(out),
                                                 _stream_ptr(dev)))
    return out
  _lib.check(_lib.lib().det_lookup_sparse(table.handle, _ptr(ids), _ptr(seg), _ptr(w), ids.numel(), batch,
                                          _lib.COMBINERS[combiner], _ptr(default_row), _ptr(out),
                                          _stream_ptr(dev)))
  return out


class _SparseSegmentSum(torch.autograd.Function):
  """gather(rows, idx) * weights -> segment sum -> normalise over a DENSE [U, dim] matrix (det_sparse_segment_sum: the
  kernels of the fused forward; the ids of a segment are added in order, so the result is deterministic and
  bit-identical to the oracle).  Differentiable w.r.t. rows: d rows[u] = sum over the ids i of u of
  gout[seg_i] * w_i / norm[seg_i], summed by `combine_rows` (the reference's gradient dedupe)."""

  @staticmethod
  def forward(ctx, rows, idx, seg, weights, batch, combiner):
    dev = rows.device
    rows_c = rows.contiguous()
    idx64 = idx.reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    seg32 = seg.reshape(-1).to(device=dev, dtype=torch.int32).contiguous()
    w = None if weights is None else weights.reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
    dim = rows_c.shape[1]
    out = torch.empty((batch, dim), dtype=torch.float32, device=dev)
    lib = _lib.lib()
    ws_bytes = lib.det_sparse_segment_sum_workspace_bytes(batch)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    zero_row = torch.zeros(dim, dtype=torch.float32, device=dev)
    _lib.check(lib.det_sparse_segment_sum(_ptr(rows_c), dim, _ptr(idx64), _ptr(seg32), _ptr(w), idx64.numel(), batch,
                                          _lib.COMBINERS[combiner], _ptr(zero_row), _ptr(out), _ptr(ws), ws_bytes,
                                          _stream_ptr(dev)))
    ctx.save_for_backward(idx64, seg32, w if w is not None else torch.empty(0, device=dev))
    ctx.meta = (rows_c.shape[0], batch, combiner, w is not None)
    return out

  @staticmethod
  def backward(ctx, gout):
    from .variable import combine_rows
    idx64, seg32, w = ctx.saved_tensors
    n_rows, batch, combiner, has_w = ctx.meta
    seg = seg32.long()
    coef = w if has_w else torch.ones(seg.numel(), dtype=torch.float32, device=gout.device)
    if combiner != "sum":
      den = torch.zeros(batch, dtype=torch.float32, device=gout.device).index_add_(0, seg, coef if combiner == "mean" else coef * coef)
      if combiner == "sqrtn":
        den = den.sqrt()
      coef = coef / den[seg]
    g_rows = gout.to(torch.float32)[seg] * coef[:, None]
    return combine_rows(g_rows.contiguous(), idx64.to(torch.int32), n_rows), None, None, None, None, None


def sparse_segment_sum_rows(rows, idx, segment_ids, weights, batch, combiner):
  """[U, dim] fp32 rows, idx [nnz] into them, ascending segment_ids [nnz] -> [batch, dim]; differentiable w.r.t. rows"""
  return _SparseSegmentSum.apply(rows, idx, segment_ids, weights, int(batch), combiner)


def embedding_lookup_sparse(params, sp_ids, sp_weights, partition_strategy=None, name="embedding_lookup_sparse",
                            combiner="mean", max_norm=None, return_trainable=False):
  """Dynamic version of tf.nn.embedding_lookup_sparse (dynamic_embedding_ops.py:120-293): for every row of
  the dense matrix represented by sp_ids, combine (sum / mean / sqrtn) the weighted embeddings of its ids.
  Result: [dense_shape[0], dim] float32.

  Forward-only calls on a single-shard fp32 variable with a constant default row run the fused kernel (with
  `max_norm` folded into the gather for rows of up to 32 vectors);
  `return_trainable=True`, sharded variables and random initializers take the composed
  path (unique -> lookup -> gather*weights -> segment sum), exactly the reference's op sequence."""
  if combiner not in ("mean", "sqrtn", "sum"):
    raise ValueError("combiner must be one of 'mean', 'sqrtn' or 'sum'")
  if not isinstance(sp_ids, SparseIds):
    raise TypeError("sp_ids must be SparseTensor")
  ignore_weights = sp_weights is None
  if not ignore_weights and not isinstance(sp_weights, SparseIds):
    raise TypeError("sp_weights must be either None or SparseTensor")
  if not isinstance(params, Variable):
    raise TypeError("params should be a Variable instance.")
  if params.key_dtype != sp_ids.values.dtype:   # raised by the inner embedding_lookup, dynamic_embedding_variable.py:1409-1412
    raise TypeError("params.key_dtype should be same with sp_ids.dtype: {} vs. {}".format(params.key_dtype,
                                                                                           sp_ids.values.dtype))
  segment_ids = sp_ids.indices[:, 0].to(torch.int32)
  ids = sp_ids.values
  batch = sp_ids.dense_shape[0]
  weights = None if ignore_weights else sp_weights.values
  static_default = params.initializer is None or not callable(params.initializer)
  if not return_trainable and static_default and _fused_ok(params, params.tables[0]._default_value) and \
      (max_norm is None or (float(max_norm) > 0 and _clip_fusable(params.dim))):
    return lookup_sparse_fused(params, ids, segment_ids, weights, batch, combiner, max_norm=max_norm)

  uniq, idx = unique(ids)
  r = embedding_lookup(params, uniq, max_norm=max_norm, return_trainable=return_trainable)
  emb_u, tw = r if return_trainable else (r, None)
  import os
  if os.environ.get("DET_SPARSE_TRAIN_FUSED", "1") == "1" and emb_u.device.type in _table_device_types() and ids.numel() > 0:
    # DEFAULT: the dense rows of the trainable scratch go through the fused gather / weight / segment-sum kernel
    # (det_sparse_segment_sum; validated on B200 in round 2, bit-identical to the oracle; its backward is the
    # position-order gradient dedupe).  DET_SPARSE_TRAIN_FUSED=0 selects the eager torch restatement below.
    out = sparse_segment_sum_rows(emb_u.to(torch.float32), idx, segment_ids, weights, batch, combiner)
    return (out, tw) if return_trainable else out
  emb = gather_unique(emb_u.to(torch.float32), idx)
  seg64 = segment_ids.long().to(emb.device)
  w = torch.ones(ids.numel(), dtype=torch.float32, device=emb.device) if ignore_weights else \
      weights.to(device=emb.device, dtype=torch.float32)
  out = torch.zeros((batch, params.dim), dtype=torch.float32, device=emb.device).index_add(0, seg64, emb * w[:, None])
  if combiner != "sum":
    den = torch.zeros(batch, dtype=torch.float32, device=emb.device).index_add(
        0, seg64, w if combiner == "mean" else w * w)
    if combiner == "sqrtn":
      den = den.sqrt()
    touched = torch.zeros(batch, dtype=torch.bool, device=emb.device)
    touched[seg64] = True
    out = torch.where(touched[:, None], out / den[:, None], out)
  return (out, tw) if return_trainable else out


def verify_embedding_param_weights(embedding_weights, sparse_ids, sparse_weights=None):
  """EmbeddingWeights.verify_embedding_param_weights (python/ops/embedding_weights.py:78-95)"""
  if embedding_weights is None:
    raise ValueError("Missing embedding_weights %s." % embedding_weights)
  if embedding_weights.key_dtype != sparse_ids.values.dtype:
    raise TypeError("embedding_weights.key_dtype should be same with sparse_ids.dtype: {} vs. {}".format(
        embedding_weights.key_dtype, sparse_ids.values.dtype))
  weights_dtype = sparse_weights.values.dtype if sparse_weights is not None else None
  if weights_dtype and embedding_weights.value_dtype != weights_dtype:
    raise TypeError("embedding_weights.value_dtype should be same with sparse_weights.dtype: {} vs. {}".format(
        embedding_weights.value_dtype, weights_dtype))


def _safe_preprocess(sparse_ids, sparse_weights, combiner, default_id):
  """The tensor surgery of safe_embedding_lookup_sparse (dynamic_embedding_ops.py:349-383), device-agnostic:
  flatten the leading dims to one row axis, prune entries with weight <= 0 (unless combiner is "sum"), give every
  empty row one (default_id or 0, weight 1) entry, keep canonical row-major order.
  Returns (SparseIds 2-D, weights SparseIds or None, indices of the originally empty rows, original dense_shape)."""
  shape = sparse_ids.dense_shape
  rank = len(shape)
  lead = 1
  for d in shape[:-1]:
    lead *= d
  idx = sparse_ids.indices
  strides = []
  s = 1
  for d in reversed(shape[:-1]):
    strides.append(s)
    s *= d
  strides = list(reversed(strides))
  row = torch.zeros(idx.shape[0], dtype=torch.int64, device=idx.device)
  for k in range(rank - 1):
    row = row + idx[:, k] * strides[k]
  ids = sparse_ids.values
  w = None if sparse_weights is None else sparse_weights.values
  col = idx[:, rank - 1]
  if combiner != "sum" and w is not None:
    keep = w > 0
    row, col, ids, w = row[keep], col[keep], ids[keep], w[keep]
  # sparse_fill_empty_rows
  present = torch.zeros(lead, dtype=torch.bool, device=idx.device)
  present[row] = True
  empty_rows = torch.nonzero(~present).reshape(-1)
  if empty_rows.numel():
    row = torch.cat([row, empty_rows])
    col = torch.cat([col, torch.zeros_like(empty_rows)])
    ids = torch.cat([ids, torch.full_like(empty_rows, default_id or 0)])
    if w is not None:
      w = torch.cat([w, torch.ones(empty_rows.numel(), dtype=w.dtype, device=w.device)])
    order = torch.sort(row * (shape[-1] + 1) + col, stable=True).indices
    row, col, ids = row[order], col[order], ids[order]
    if w is not None:
      w = w[order]
  ind2 = torch.stack([row, col], 1)
  sp2 = SparseIds(ind2, ids, (lead, shape[-1]))
  sw2 = None if w is None else SparseIds(ind2, w, (lead, shape[-1]))
  return sp2, sw2, empty_rows, shape


def safe_embedding_lookup_sparse(embedding_weights, sparse_ids, sparse_weights=None, combiner="mean",
                                 default_id=None, name="safe_embedding_lookup_sparse", partition_strategy=None,
                                 max_norm=None, return_trainable=False):
  """dynamic_embedding_ops.py:296-438: flatten leading dims, prune weights <= 0 (unless combiner is "sum"),
  give empty rows `default_id` (or a zero vector when default_id is None), then embedding_lookup_sparse."""
  verify_embedding_param_weights(embedding_weights, sparse_ids, sparse_weights)
  sp2, sw2, empty_rows, shape = _safe_preprocess(sparse_ids, sparse_weights, combiner, default_id)
  r = embedding_lookup_sparse(embedding_weights, sp2, sw2, combiner=combiner, max_norm=max_norm,
                              return_trainable=return_trainable)
  result, tw = r if return_trainable else (r, None)
  if default_id is None and empty_rows.numel():
    result = result.clone()
    result[empty_rows] = 0
  final = result.reshape(tuple(shape[:-1]) + (result.shape[-1],))
  return (final, tw) if return_trainable else final
