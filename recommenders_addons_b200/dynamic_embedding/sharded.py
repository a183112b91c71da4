"""Key-hash sharding of one logical table across the GPUs of a box: one process per GPU, each owning the
shard `owner(key) == rank` (reference: HvdVariable.__alltoall_embedding_lookup__,
python/ops/shadow_embedding_ops.py:397-447, with Horovod alltoall replaced by NCCL through
torch.distributed).  Exchange per lookup: all-to-all of keys (8 B/key) -> local find -> all-to-all of rows
back (dim*4 B/key); per update: all-to-all of keys + rows/grads to the owners -> local insert / fused
optimizer.  Works with the `gloo` backend on CPU tensors for the exchange logic tests only when a
`local_table` stub is injected; the real table is always on the GPU."""
import torch
import torch.distributed as dist

from .variable import Variable, gather_rows, partition, scatter_rows


def _alltoall_counts(counts, group):
  """send counts [W] -> recv counts [W] (the 8x8 size exchange)."""
  recv = torch.empty_like(counts)
  dist.all_to_all_single(recv, counts, group=group)
  return recv


def exchange_keys(grouped_keys, send_counts, group=None):
  """all-to-all(v) of keys grouped by owner.  Returns (received keys, recv_counts list)."""
  recv_counts = _alltoall_counts(send_counts, group)
  sc, rc = send_counts.tolist(), recv_counts.tolist()
  out = torch.empty(int(sum(rc)), dtype=grouped_keys.dtype, device=grouped_keys.device)
  dist.all_to_all_single(out, grouped_keys, output_split_sizes=rc, input_split_sizes=sc, group=group)
  return out, rc, sc


def exchange_rows(rows, in_splits, out_splits, group=None):
  out = torch.empty((int(sum(out_splits)),) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
  dist.all_to_all_single(out, rows.contiguous(), output_split_sizes=out_splits, input_split_sizes=in_splits,
                         group=group)
  return out


class ShardedVariable(object):
  """One shard of a key-hash-sharded variable per rank; lookups / updates take keys owned by ANY rank."""

  def __init__(self, local_variable, group=None, partition_impl=None, gather_impl=None, scatter_impl=None):
    self.local = local_variable
    self.group = group
    self.world = dist.get_world_size(group)
    self.rank = dist.get_rank(group)
    self._partition = partition_impl or (lambda k: partition(k, self.world, True))
    self._gather = gather_impl or gather_rows
    self._scatter = scatter_impl or scatter_rows

  @property
  def dim(self):
    return self.local.dim

  def lookup(self, keys):
    """rows for `keys` [n] in request order, wherever they live."""
    grouped, perm, counts = self._partition(keys.reshape(-1))
    recv_keys, rc, sc = exchange_keys(grouped, counts, self.group)
    rows = self.local.lookup(recv_keys)                      # local shard: find (no insert)
    back = exchange_rows(rows, rc, sc, self.group)          # rows return on the reverse pattern
    return self._scatter(back, perm)

  def upsert(self, keys, values):
    grouped, perm, counts = self._partition(keys.reshape(-1))
    vals = self._gather(values.reshape(-1, self.dim), perm)
    recv_keys, rc, sc = exchange_keys(grouped, counts, self.group)
    recv_vals = exchange_rows(vals, sc, rc, self.group)
    self.local.upsert(recv_keys, recv_vals)

  def apply_gradients(self, optimizer, keys, grads):
    """route row-gradients to the owning rank, which applies the fused optimizer locally (half-sync:
    sparse rows are never all-reduced, dynamic_embedding_optimizer.py:580-595).  Keys that several ranks
    send for the same row are combined (summed) on the owner before the update."""
    grouped, perm, counts = self._partition(keys.reshape(-1))
    g = self._gather(grads.reshape(-1, self.dim), perm)
    recv_keys, rc, sc = exchange_keys(grouped, counts, self.group)
    recv_g = exchange_rows(g, sc, rc, self.group)
    from .variable import unique
    uniq, idx = unique(recv_keys)
    gsum = torch.zeros((uniq.numel(), self.dim), dtype=recv_g.dtype, device=recv_g.device).index_add(
        0, idx.long(), recv_g)
    optimizer.iterations += 1
    optimizer.apply_sparse(self.local, uniq, gsum)

  def size(self):
    s = self.local.size().to(torch.int64)
    t = s.clone().to(next(iter(self.local.tables)).device) if hasattr(self.local, "tables") else s.clone()
    dist.all_reduce(t, group=self.group)
    return t
