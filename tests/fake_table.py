"""A dict-backed stand-in for the table object (LookupInterface subset `de.Variable` calls), TEST-ONLY: it lets the
host-side logic that is pure composition of table ops (restrict policies, Variable plumbing) be checked without a GPU.
The product never uses it: `de.CuckooHashTable` / `de.HkvHashTable` have no CPU fallback."""
import torch

from recommenders_addons_b200.dynamic_embedding.table import KVCreator


class DictTable(object):

  def __init__(self, key_dtype, value_dtype, default_value, name, num_slot_planes=0):
    self._default_value = torch.as_tensor(default_value, dtype=value_dtype).reshape(-1)
    self._dim = int(self._default_value.numel())
    self._value_dtype = value_dtype
    self._name = name
    self._num_slot_planes = num_slot_planes
    self.device = torch.device("cpu")
    self.d = {}

  def size(self, name=None):
    return torch.tensor(len(self.d), dtype=torch.int64)

  def lookup(self, keys, dynamic_default_values=None, return_exists=False, name=None):
    flat = keys.reshape(-1).tolist()
    default = self._default_value if dynamic_default_values is None else torch.as_tensor(
        dynamic_default_values, dtype=self._value_dtype).reshape(-1)
    full = default.numel() == len(flat) * self._dim and len(flat) > 0
    rows, ex = [], []
    for i, k in enumerate(flat):
      if k in self.d:
        rows.append(self.d[k])
        ex.append(True)
      else:
        rows.append(default[i * self._dim:(i + 1) * self._dim] if full else default[:self._dim])
        ex.append(False)
    vals = torch.stack(rows).reshape(tuple(keys.shape) + (self._dim,)) if rows else torch.empty(
        tuple(keys.shape) + (self._dim,), dtype=self._value_dtype)
    if return_exists:
      return vals, torch.tensor(ex, dtype=torch.bool).reshape(keys.shape)
    return vals

  def insert(self, keys, values, name=None):
    values = values.reshape(-1, self._dim)
    for i, k in enumerate(keys.reshape(-1).tolist()):
      self.d[k] = values[i].clone()

  def remove(self, keys, name=None):
    for k in keys.reshape(-1).tolist():
      self.d.pop(k, None)

  def clear(self, name=None):
    self.d.clear()

  def export(self, name=None, plane=0):
    ks = list(self.d.keys())
    keys = torch.tensor(ks, dtype=torch.int64)
    vals = torch.stack([self.d[k] for k in ks]) if ks else torch.empty((0, self._dim), dtype=self._value_dtype)
    return keys, vals


class DictTableCreator(KVCreator):

  def create(self, key_dtype=None, value_dtype=None, default_value=None, name=None, checkpoint=None, init_size=None,
             config=None, device=None, shard_saveable_object_fn=None, num_slot_planes=0):
    return DictTable(key_dtype, value_dtype, default_value, name, num_slot_planes)
