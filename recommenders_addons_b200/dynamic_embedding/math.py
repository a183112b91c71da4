"""Mirror of `tfra.dynamic_embedding.math` (python/ops/math_ops.py:60-215): the sparse helpers the lookup path is built
from in the reference -- `sparse_segment_sum` (kernel `SortedSparseSegmentSumCustomKernel`,
core/kernels/segment_reduction_ops_gpu.cu.cc:29-113), `sparse_fill_empty_rows`, `sparse_reshape`.  The fused forward
(`det_lookup_sparse`) never materialises these intermediates; the functions exist for callers that use them directly.
Device-agnostic torch code (plumbing); rows are accumulated in index order, like the reference's CPU kernel."""
import torch

from .ops import SparseIds


def sparse_segment_sum(data, indices, segment_ids, name=None, num_segments=None):
  """output[s] = sum over i with segment_ids[i] == s of data[indices[i]]; segment_ids sorted, may repeat (:60-137)"""
  indices = torch.as_tensor(indices, device=data.device).reshape(-1).long()
  segment_ids = torch.as_tensor(segment_ids, device=data.device).reshape(-1).long()
  if indices.numel() != segment_ids.numel():
    raise ValueError("indices and segment_ids should have the same length")
  if segment_ids.numel() > 1 and bool((segment_ids[1:] < segment_ids[:-1]).any()):
    raise ValueError("segment ids are not increasing")
  k = int(num_segments) if num_segments is not None else (int(segment_ids[-1]) + 1 if segment_ids.numel() else 0)
  out = torch.zeros((k,) + tuple(data.shape[1:]), dtype=data.dtype, device=data.device)
  if segment_ids.numel():
    out.index_add_(0, segment_ids, data[indices])
  return out


def sparse_fill_empty_rows(sp_input, default_value, name=None):
  """tf.sparse.fill_empty_rows (:168-189): every row of the 2-D sparse input without entries gets (row, 0) ->
  default_value.  Returns (SparseIds in canonical row-major order, bool[rows] "was empty")."""
  rows, cols = int(sp_input.dense_shape[0]), int(sp_input.dense_shape[1])
  idx, vals = sp_input.indices, sp_input.values
  present = torch.zeros(rows, dtype=torch.bool, device=idx.device)
  if idx.numel():
    present[idx[:, 0]] = True
  empty = torch.nonzero(~present).reshape(-1)
  if empty.numel():
    add = torch.stack([empty, torch.zeros_like(empty)], 1)
    idx = torch.cat([idx, add], 0)
    vals = torch.cat([vals, torch.full((empty.numel(),), default_value, dtype=vals.dtype, device=vals.device)], 0)
    order = torch.sort(idx[:, 0] * (cols + 1) + idx[:, 1], stable=True).indices
    idx, vals = idx[order], vals[order]
  return SparseIds(idx, vals, (rows, cols)), ~present


def sparse_reshape(sp_input, shape, name=None):
  """tf.sparse.reshape (:192-215): same entries, indices re-expressed in the new dense shape (one -1 is inferred)"""
  old = [int(d) for d in sp_input.dense_shape]
  total = 1
  for d in old:
    total *= d
  shape = [int(d) for d in shape]
  if shape.count(-1) > 1:
    raise ValueError("only one dimension of the new shape may be -1")
  if -1 in shape:
    known = 1
    for d in shape:
      if d != -1:
        known *= d
    if known == 0 or total % known:
      raise ValueError("cannot infer the -1 dimension: %s -> %s" % (old, shape))
    shape[shape.index(-1)] = total // known
  new_total = 1
  for d in shape:
    new_total *= d
  if new_total != total:
    raise ValueError("Input to reshape is a SparseTensor with %d dense values, but the requested shape has %d" %
                     (total, new_total))
  idx = sp_input.indices
  flat = torch.zeros(idx.shape[0], dtype=torch.int64, device=idx.device)
  for k, d in enumerate(old):
    flat = flat * d + idx[:, k]
  cols = []
  for d in reversed(shape):
    cols.append(flat % d)
    flat = flat // d
  new_idx = torch.stack(list(reversed(cols)), 1) if cols else idx.new_zeros((idx.shape[0], 0))
  return SparseIds(new_idx, sp_input.values, tuple(shape))
