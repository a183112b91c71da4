"""Mirror of `tfra.dynamic_embedding.shadow_ops` (python/ops/shadow_embedding_ops.py:57-330): `ShadowVariable`, the
persistent eager twin of TrainableWrapper that the Keras layers look rows up through, and the shadow flavours of
`embedding_lookup` / `embedding_lookup_unique`.  A ShadowVariable projects the activated part of the sparse domain
(`params`, a de.Variable) onto a dense scratch: `embedding_lookup(shadow, ids)` stores the ids, fills the scratch from
the table and returns it; the optimizer steps the scratch and `update_op` writes it back."""
import torch

from .variable import ModelMode, TrainableWrapper, Variable, unique


class ShadowVariable(TrainableWrapper):
  """shadow_embedding_ops.py:57-187"""

  def __init__(self, params, name="ShadowVariable", max_norm=None, trainable=True, distribute_strategy=None, **kwargs):
    if not isinstance(params, Variable):
      raise TypeError("params must be de.Variable, but get %s" % type(params))
    ids = kwargs.get("ids", None)
    if ids is None:
      ids = torch.zeros((0,), dtype=params.key_dtype, device=params.tables[0].device)
    elif not torch.is_tensor(ids):
      raise TypeError("If ids is set, it needs to be a tensor buffer")
    super().__init__(params, ids, values=None, exists=kwargs.get("exists", None), max_norm=max_norm,
                     model_mode=kwargs.get("model_mode", None))
    self.name = name
    self.trainable = trainable
    params.trainable_store[name] = self         # :163

  def prefetch_values(self, update=False):
    out = super().prefetch_values(update=update)
    if not self.trainable and self.values is not None and self.values.requires_grad:
      self.values = self.values.detach()
      out = self.transform(self.values)
    return out

  def embedding_lookup(self, ids, name=None, max_norm=None, return_trainable=False):
    """:168-176"""
    if return_trainable:
      return embedding_lookup(self, ids, name=name), self
    return embedding_lookup(self, ids, name=name)

  def value(self, do_prefetch=False):
    """:189-195"""
    return self.read_value(do_prefetch=do_prefetch)

  def verify_embedding_weights(self, sparse_ids, sparse_weights=None):
    """:166-168"""
    self.params.verify_embedding_weights(sparse_ids, sparse_weights)

  def assign(self, value, use_locking=None, name=None, read_value=True):
    """:198-226: a new value for the shadow's lookup buffer (one row per current id); `update_op` writes it back"""
    value = torch.as_tensor(value, dtype=self.params.value_dtype, device=self.params.tables[0].device)
    if value.numel() != self.ids.numel() * self.params.dim:
      raise ValueError("assign: expected %d rows of %d elements, got %s" % (self.ids.numel(), self.params.dim, tuple(value.shape)))
    value = value.reshape(self.ids.numel(), self.params.dim).detach().clone()
    self.values = value.requires_grad_(True) if (self.trainable and value.is_floating_point()) else value
    return self.values if read_value else None

  def _reset_ids(self, ids):
    """:231-232"""
    self.ids = ids
    return self.ids


def embedding_lookup(shadow, ids, partition_strategy=None, name=None, validate_indices=None):
  """shadow_embedding_ops.py:242-282: ids of any shape -> ids.shape + [dim]; in TRAIN mode the ids are kept and the
  dense scratch refreshed (`read_value(do_prefetch=True)`), in INFERENCE mode a plain `params.lookup`."""
  if not isinstance(shadow, ShadowVariable):
    raise TypeError("shadow must be a ShadowVariable")
  if not torch.is_tensor(ids):
    ids = torch.as_tensor(ids)
  if shadow.params.key_dtype != ids.dtype:
    raise ValueError("{} ids is not matched with ShadowVariable with ids {},".format(ids.dtype, shadow.params.key_dtype))
  shape = tuple(ids.shape) + (shadow.params.dim,)
  if ModelMode.CURRENT_SETTING == ModelMode.TRAIN:
    shadow._reset_ids(ids.reshape(-1))
    return shadow.read_value(do_prefetch=True).reshape(shape)
  return shadow.params.lookup(ids.reshape(-1)).reshape(shape)


def embedding_lookup_unique_base(ids, embedding_size, lookup_function, with_unique=True, name=None):
  """:285-330: optional unique -> lookup_function(unique ids) -> gather back, result ids.shape + [embedding_size]"""
  if not torch.is_tensor(ids):
    ids = torch.as_tensor(ids)
  shape = tuple(ids.shape) + (int(embedding_size),)
  flat = ids.reshape(-1)
  if with_unique:
    uniq, idx = unique(flat)
    emb = lookup_function(uniq)[idx.long()]
  else:
    emb = lookup_function(flat)
  return emb.reshape(shape)


def embedding_lookup_unique(shadow, ids, embedding_size, with_unique=True, name=None):
  """:333-350"""
  return embedding_lookup_unique_base(ids, embedding_size, lambda x: embedding_lookup(shadow, x), with_unique, name)
