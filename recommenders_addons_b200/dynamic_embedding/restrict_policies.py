"""Restrict policies (reference: python/ops/restrict_policies.py:36-361): rules that keep track of a per-key status
(last-seen timestamp / frequency) in a side `de.Variable` and shrink the target variable to `num_reserved` keys by
removing the oldest / least frequent ones.  Same classes, methods and errors as the reference; every step is a table
op of the engine (lookup / upsert / export / remove) plus a top-k over the exported status.

Optimizer slots live in planes of the SAME table here, so removing a key from the variable removes its
slot rows with it: `_track_params_from_optimizer_slots` only records what it is given, for API compatibility.
(Tables created with an eviction strategy do the same job inside the engine: HkvHashTable.evict / det_evict.)"""
import time

import torch


class RestrictPolicy(object):
  """restrict_policies.py:36-112"""

  def __init__(self, var):
    self.var = var
    self.params_in_slots = []

  def apply_update(self, ids):
    raise NotImplementedError

  def apply_restriction(self, num_reserved, **kwargs):
    raise NotImplementedError

  @property
  def status(self):
    return None

  def _track_params_from_optimizer_slots(self, slots):
    from .variable import TrainableWrapper, Variable
    for _s in slots:
      if isinstance(_s, TrainableWrapper):
        params = _s.params
      elif isinstance(_s, Variable):
        params = _s
      else:
        raise TypeError("slots should be dynamic_embedding.TrainableWrapper"
                        "or dynamic_embedding.Variable. But get {}".format(type(_s)))
      if id(params) not in [id(p) for p in self.params_in_slots]:
        self.params_in_slots.append(params)

  # shared by the two policies below (restrict_policies.py:196-231, 312-357)
  def _status_var(self, suffix):
    from .variable import get_variable
    return get_variable(self.var.name + suffix, key_dtype=self.var.key_dtype, value_dtype=torch.int32, dim=1,
                        devices=self.var.devices, partitioner=self.var.partition_fn, trainable=False,
                        init_size=self.var.init_size, kv_creator=self.var.kv_creator)

  def _check_restriction_args(self, num_reserved, kwargs):
    if not isinstance(num_reserved, int):
      raise TypeError("num_reserved should be integer.")
    if num_reserved < 0:
      raise ValueError("num_reserved should be non-negative.")
    trigger = kwargs.get("trigger", num_reserved)
    if not isinstance(trigger, int):
      raise TypeError("trigger should be integer.")
    return trigger

  def _restrict_lowest(self, status_var, num_reserved):
    """per shard: export the status, take the (n - reserved) LOWEST values (top_k of the negated status,
    :213-222), remove those keys from the variable, the status variable and the tracked slot variables"""
    for i in range(status_var.shard_num):
      keys, stat = status_var.tables[i].export()
      stat = stat.reshape(-1)
      reserved = int(num_reserved / status_var.shard_num)
      k = max(int(stat.numel()) - reserved, 0)
      if k == 0:
        continue
      idx = torch.topk(-stat.to(torch.int64), k, sorted=False).indices
      removed = keys[idx].contiguous()
      self.var.tables[i].remove(removed)
      status_var.tables[i].remove(removed)
      for slot_param in self.params_in_slots:
        slot_param.tables[i].remove(removed)


class TimestampRestrictPolicy(RestrictPolicy):
  """`oldest-out-first` (restrict_policies.py:115-235): status = int32 timestamp (seconds) of the last update"""

  def __init__(self, var):
    super().__init__(var)
    self.tstp_var = self._status_var("/timestamp")
    self._restrict_var = self.tstp_var

  def apply_update(self, ids):
    keys = torch.unique(ids.reshape(-1))
    fresh = torch.full((keys.numel(), 1), int(time.time()) & 0x7fffffff, dtype=torch.int32, device=keys.device)
    self.tstp_var.upsert(keys, fresh)

  def apply_restriction(self, num_reserved, **kwargs):
    trigger = self._check_restriction_args(num_reserved, kwargs)
    self._num_reserved = num_reserved
    if int(self.var.size()) > trigger:
      self._restrict_lowest(self.tstp_var, num_reserved)

  @property
  def status(self):
    return self.tstp_var


class FrequencyRestrictPolicy(RestrictPolicy):
  """`lowest-occurrence-out-first` (restrict_policies.py:238-361): status = int32 number of updates seen"""

  def __init__(self, var):
    super().__init__(var)
    self.init_count = 0
    self.freq_var = self._status_var("/frequency")
    self._restrict_var = self.freq_var

  def apply_update(self, ids):
    keys = torch.unique(ids.reshape(-1))
    counts = self.freq_var.lookup(keys)      # absent keys read the initializer, 0
    self.freq_var.upsert(keys, counts + 1)

  def apply_restriction(self, num_reserved, **kwargs):
    trigger = self._check_restriction_args(num_reserved, kwargs)
    self._num_reserved = num_reserved
    if int(self.var.size()) > trigger:
      self._restrict_lowest(self.freq_var, num_reserved)

  @property
  def status(self):
    return self.freq_var
