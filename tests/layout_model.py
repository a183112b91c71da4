"""Sequential NumPy/Python model of the GPU table's LAYOUT LOGIC (recommenders_addons_b200/csrc/common.cuh):
8-slot buckets, home bucket = mulhi64(fmix64(key), nb), linear probing over buckets, chain ends at the first bucket
that still holds an EMPTY slot, insert = first free (EMPTY or TOMBSTONE) slot of the chain, erase = EMPTY if the
key's bucket still has an EMPTY slot else TOMBSTONE, the two sentinel key values served by side slots.
It exists to property-test the invariants the CUDA kernels rely on, without a GPU (tests/test_layout_model.py)."""
import numpy as np

BUCKET = 8
EMPTY = -(1 << 63)
TOMB = EMPTY + 1
M64 = (1 << 64) - 1


def fmix64(k):
  k &= M64
  k ^= k >> 33
  k = (k * 0xff51afd7ed558ccd) & M64
  k ^= k >> 33
  k = (k * 0xc4ceb9fe1a85ec53) & M64
  k ^= k >> 33
  return k


class LayoutModel(object):

  def __init__(self, nb):
    self.nb = nb
    self.keys = np.full(nb * BUCKET, EMPTY, dtype=object)
    self.vals = {}
    self.special = [False, False]
    self.size = 0
    self.used = 0

  def home(self, key):
    return (fmix64(int(key)) * self.nb) >> 64

  def _chain(self, key):
    """yields (bucket index, slot range) along the probe chain until (and including) the bucket that ends it"""
    b = self.home(key)
    for _ in range(self.nb):
      lo = b * BUCKET
      yield b, range(lo, lo + BUCKET)
      if any(self.keys[s] == EMPTY for s in range(lo, lo + BUCKET)):
        return
      b = (b + 1) % self.nb

  def find(self, key):
    if key in (EMPTY, TOMB):
      i = int(key == TOMB)
      return self.nb * BUCKET + i if self.special[i] else -1
    for _, slots in self._chain(key):
      for s in slots:
        if self.keys[s] == key:
          return s
    return -1

  def insert(self, key, value):
    if key in (EMPTY, TOMB):
      i = int(key == TOMB)
      if not self.special[i]:
        self.special[i] = True
        self.size += 1
      self.vals[self.nb * BUCKET + i] = value
      return
    first_free = -1
    for _, slots in self._chain(key):
      for s in slots:
        if self.keys[s] == key:
          self.vals[s] = value
          return
        if first_free < 0 and self.keys[s] in (EMPTY, TOMB):
          first_free = s
    assert first_free >= 0, "table full"
    if self.keys[first_free] == EMPTY:
      self.used += 1
    self.keys[first_free] = key
    self.vals[first_free] = value
    self.size += 1

  def remove(self, key):
    s = self.find(key)
    if s < 0:
      return
    if s >= self.nb * BUCKET:
      self.special[s - self.nb * BUCKET] = False
      self.size -= 1
      return
    lo = (s // BUCKET) * BUCKET
    has_empty = any(self.keys[q] == EMPTY for q in range(lo, lo + BUCKET))
    self.keys[s] = EMPTY if has_empty else TOMB
    if has_empty:
      self.used -= 1
    self.size -= 1

  def live(self):
    out = {self.keys[s]: self.vals[s] for s in range(self.nb * BUCKET) if self.keys[s] not in (EMPTY, TOMB)}
    for i, k in enumerate((EMPTY, TOMB)):
      if self.special[i]:
        out[k] = self.vals[self.nb * BUCKET + i]
    return out

  def check_invariants(self):
    ks = [k for k in self.keys if k not in (EMPTY, TOMB)]
    assert len(ks) == len(set(ks)), "a key is stored twice"
    assert self.size == len(ks) + sum(self.special)
    assert self.used == sum(1 for k in self.keys if k != EMPTY)
    for s, k in enumerate(self.keys):
      if k in (EMPTY, TOMB):
        continue
      assert self.find(k) == s, "a stored key is not reachable along its probe chain"
