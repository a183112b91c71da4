"""Mirror of `tfra.dynamic_embedding.data_flow` (python/ops/data_flow_ops.py:40-61; GPU kernels
core/kernels/dynamic_partition_op_gpu.cu.cc, dynamic_stitch_op_gpu.cu.cc): what `Variable` uses to split keys / rows by
shard and to merge the per-shard results (`make_partition` / `_stitch`, dynamic_embedding_variable.py:131-163).
Inside `de.Variable` the same job is ONE counting-partition kernel (`det_partition`) plus row permutes
(`det_gather_rows` / `det_scatter_rows`); these functions give the op-level semantics to callers that use them directly."""
import torch


def dynamic_partition(data, partitions, num_partitions, name=None):
  """tf.dynamic_partition: outputs[p] = the slices data[i] with partitions[i] == p, in their original order"""
  partitions = torch.as_tensor(partitions, device=data.device)
  if tuple(data.shape[:partitions.dim()]) != tuple(partitions.shape):
    raise ValueError("partitions.shape must be a prefix of data.shape")
  flat_p = partitions.reshape(-1)
  if flat_p.numel() and (int(flat_p.min()) < 0 or int(flat_p.max()) >= int(num_partitions)):
    raise ValueError("partitions must be in [0, %d)" % int(num_partitions))
  flat = data.reshape((flat_p.numel(),) + tuple(data.shape[partitions.dim():]))
  return [flat[flat_p == p] for p in range(int(num_partitions))]


def dynamic_stitch(indices, data, use_fast=True, name=None):
  """tf.dynamic_stitch: merged[indices[m][i]] = data[m][i]; later entries win when an index repeats"""
  if torch.is_tensor(indices):
    indices, data = [indices], [data]
  if len(indices) != len(data):
    raise ValueError("indices and data must have the same length")
  if not indices:
    raise ValueError("dynamic_stitch needs at least one (indices, data) pair")
  n = 0
  for ind in indices:
    if ind.numel():
      n = max(n, int(ind.max()) + 1)
  inner = tuple(data[0].shape[indices[0].dim():])
  out = torch.zeros((n,) + inner, dtype=data[0].dtype, device=data[0].device)
  for ind, d in zip(indices, data):
    if tuple(d.shape[:ind.dim()]) != tuple(ind.shape):
      raise ValueError("data[m].shape must start with indices[m].shape")
    fi = ind.reshape(-1).long()
    fd = d.reshape((fi.numel(),) + inner)
    # in-order assignment (a repeated index keeps its LAST value, like the reference kernels)
    if fi.numel() and fi.unique().numel() != fi.numel():
      for j in range(fi.numel()):
        out[fi[j]] = fd[j]
    else:
      out[fi] = fd
  return out
