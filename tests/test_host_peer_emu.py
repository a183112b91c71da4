"""host_api.cu and sharded.cu on the SIMT emulator through the real C ABI:
(1) host-buffer entry points (chunked pipelines) and the `-keys` / `-values` file format of det_save / det_load
    (reference: core/kernels/cuckoo_hashtable_op.cc:310-504: raw little-endian int64 keys, raw rows);
(2) the one-sided sharded table: shards faked on one device (the reference's own test trick) and a REAL multi-rank
    group -- every rank a Python thread with its own table and peer group, peers mapped through the (emulated) CUDA-IPC
    handles -- with the flag barrier, det_peer_find / det_peer_insert against default_partition_fn, and the
    route -> barrier -> inbox exchange of the backward path.
The same code runs on 2-8 B200s in tests/test_peer_gpu.py / tests/test_multigpu_gpu.py; here it is a CPU regression net."""
import ctypes
import os
import threading

import numpy as np
import pytest

os.environ.setdefault("DET_HOST_CHUNK_MB", "1")     # several pipeline chunks at test sizes (read once by the library)

from oracle import oracle as O  # noqa: E402
from recommenders_addons_b200 import _lib as real  # noqa: E402
from tests.test_detable_emu import L, P, Table, ck  # noqa: E402

_EXTRA = ["det_find_host", "det_insert_host", "det_find_host_async", "det_insert_host_async", "det_host_sync", "det_host_sync_pipe", "det_save",
          "det_load", "det_export_window", "det_peer_handle_bytes", "det_peer_export", "det_peer_group_create", "det_peer_group_destroy",
          "det_peer_find", "det_peer_insert", "det_peer_barrier", "det_peer_inbox_bytes", "det_peer_inbox_attach",
          "det_peer_route", "det_peer_inbox_counts", "det_peer_inbox_gather", "det_table_region_bytes",
          "det_table_create_in_region", "det_peer_group_create_regions", "det_peer_xchg_bytes", "det_peer_xchg_attach",
          "det_peer_xchg_find", "det_peer_xchg_insert", "det_peer_xchg_apply_workspace_bytes", "det_peer_xchg_apply_adagrad",
          "det_peer_xchg_apply_adam"]


def X():
  l = L()
  if not getattr(l, "_extra_ready", False):
    for name in _EXTRA:
      res, args = real.SIGNATURES[name]
      fn = getattr(l, name)
      fn.restype, fn.argtypes = res, args
    l._extra_ready = True
  return l


# ---- (1) host buffers and files -------------------------------------------------------------------------------------
def test_host_buffer_entry_points_and_async_pipeline():
  rng = np.random.default_rng(0)
  dim, n = 64, 11000                                  # 1 MiB chunks of 256 B rows: 3 chunks per call
  t = Table(dim=dim, init=1 << 15)
  keys = rng.choice(1 << 40, size=n, replace=False).astype(np.int64)
  vals = rng.standard_normal((n, dim)).astype(np.float32)
  ck(X().det_insert_host(t.h, P(keys), P(vals), n))
  assert t.size() == n
  q = np.concatenate([keys[::2], np.arange(5, dtype=np.int64) - 100])
  default = rng.standard_normal((len(q), dim)).astype(np.float32)
  out = np.empty((len(q), dim), dtype=np.float32)
  ex = np.empty(len(q), dtype=np.uint8)
  ck(X().det_find_host(t.h, P(q), len(q), P(default), 1, P(out), P(ex)))
  assert ex[:-5].all() and not ex[-5:].any()
  np.testing.assert_array_equal(out[:-5], vals[::2])
  np.testing.assert_array_equal(out[-5:], default[-5:])
  # asynchronous pair: a lookup and a write-back in flight together, then one sync
  keys2 = rng.choice(1 << 40, size=4000, replace=False).astype(np.int64) + (1 << 41)
  vals2 = rng.standard_normal((4000, dim)).astype(np.float32)
  out2 = np.empty((n, dim), dtype=np.float32)
  ex2 = np.empty(n, dtype=np.uint8)
  zero = np.zeros(dim, dtype=np.float32)
  ck(X().det_find_host_async(t.h, P(keys), n, P(zero), 0, P(out2), P(ex2)))
  ck(X().det_insert_host_async(t.h, P(keys2), P(vals2), 4000))
  ck(X().det_host_sync(t.h))
  assert ex2.all()
  np.testing.assert_array_equal(out2, vals)
  assert t.size() == n + 4000
  # waiting for one pipeline only: the looked-up rows are on the host after sync_pipe(0), the write-back after (1)
  keys3 = keys2 + 7
  out3 = np.full((4000, dim), np.nan, dtype=np.float32)
  ck(X().det_find_host_async(t.h, P(keys2), 4000, P(zero), 0, P(out3), None))
  ck(X().det_insert_host_async(t.h, P(keys3), P(vals2), 4000))
  ck(X().det_host_sync_pipe(t.h, 0))
  np.testing.assert_array_equal(out3, vals2)
  ck(X().det_host_sync_pipe(t.h, 1))
  assert t.size() == n + 8000
  ck(X().det_host_sync_pipe(t.h, -1))
  assert X().det_host_sync_pipe(t.h, 2) == 1 and X().det_host_sync_pipe(None, 0) == 1
  t.close()


@pytest.mark.parametrize("dtype,dim", [(np.float32, 16), (np.int64, 3), (np.int8, 5)])
def test_save_load_file_format(tmp_path, dtype, dim):
  rng = np.random.default_rng(1)
  t = Table(dim=dim, dtype=dtype, init=64)
  keys = rng.choice(1 << 40, size=3000, replace=False).astype(np.int64)
  vals = (rng.integers(-100, 100, size=(3000, dim))).astype(dtype)
  t.insert(keys, vals)
  prefix = str(tmp_path / "tbl_mht_1of1")
  ck(X().det_save(t.h, prefix.encode(), 700, 0))            # buffer of 700 keys: several export windows
  fk = np.fromfile(prefix + "-keys", dtype="<i8")
  fv = np.fromfile(prefix + "-values", dtype=np.dtype(dtype).newbyteorder("<")).reshape(-1, dim)
  assert len(fk) == 3000 and fv.shape == (3000, dim)
  o = np.argsort(keys)
  of = np.argsort(fk)
  np.testing.assert_array_equal(fk[of], keys[o])
  np.testing.assert_array_equal(fv[of], vals[o])
  # a file pair written the way the reference writes it loads into a fresh table (load = clear + insert)
  k2 = rng.choice(1 << 40, size=1200, replace=False).astype(np.int64)
  v2 = rng.integers(-100, 100, size=(1200, dim)).astype(dtype)
  prefix2 = str(tmp_path / "other")
  k2.astype("<i8").tofile(prefix2 + "-keys")
  v2.tofile(prefix2 + "-values")
  t2 = Table(dim=dim, dtype=dtype, init=64)
  t2.insert(np.array([1, 2, 3], dtype=np.int64), np.zeros((3, dim), dtype=dtype))
  ck(X().det_load(t2.h, prefix2.encode(), 500, 1))
  assert t2.size() == 1200
  out, ex = t2.find(k2)
  assert ex.all()
  np.testing.assert_array_equal(out, v2)
  assert X().det_load(t2.h, str(tmp_path / "missing").encode(), 500, 1) == 7     # DET_IO_ERROR
  # load_entire_dir: the second file of a set is ADDED (clear_first = 0); append_to_file appends to a file pair
  ck(X().det_load(t2.h, prefix.encode(), 800, 0))
  assert t2.size() == 1200 + 3000
  ck(X().det_save(t2.h, prefix2.encode(), 512, 1))
  assert len(np.fromfile(prefix2 + "-keys", dtype="<i8")) == 1200 + 4200
  assert np.fromfile(prefix2 + "-values", dtype=dtype).size == (1200 + 4200) * dim
  # windows of the export: disjoint, in table order, together the whole table
  whole_k, whole_v = t.export()
  got_k, first = [], 0
  while True:
    kb, vb = np.empty(450, np.int64), np.empty((450, dim), dtype)
    n = ctypes.c_int64(0)
    ck(X().det_export_window(t.h, 0, first, P(kb), P(vb), 450, ctypes.byref(n), None))
    if n.value == 0:
      break
    np.testing.assert_array_equal(kb[:n.value], whole_k[first:first + n.value])
    np.testing.assert_array_equal(vb[:n.value], whole_v[first:first + n.value])
    got_k.append(kb[:n.value].copy())
    first += n.value
  assert first == 3000 and len(np.unique(np.concatenate(got_k))) == 3000
  t.close()
  t2.close()


# ---- (2) the one-sided sharded table ------------------------------------------------------------------------------
class PeerGroup(object):

  def __init__(self, tables, handles, world, rank, gpu_mode=True):
    arr = (ctypes.c_void_p * world)(*[(t.h if t is not None else None) for t in tables])
    self.g = ctypes.c_void_p()
    self._keep = handles
    ck(X().det_peer_group_create(ctypes.byref(self.g), arr, handles, world, rank, 1 if gpu_mode else 0))
    self.world, self.rank = world, rank

  def find(self, keys, dim, default=None):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    default = np.zeros(dim, dtype=np.float32) if default is None else default
    out = np.empty((len(keys), dim), dtype=np.float32)
    ex = np.empty(len(keys), dtype=np.uint8)
    ck(X().det_peer_find(self.g, P(keys), len(keys), P(default), 0, P(out), P(ex), None))
    return out, ex.astype(bool)

  def insert(self, keys, vals):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    ck(X().det_peer_insert(self.g, P(keys), P(vals), len(keys), None))

  def barrier(self):
    ck(X().det_peer_barrier(self.g, None))

  def close(self):
    if self.g:
      X().det_peer_group_destroy(self.g)
      self.g = None


@pytest.mark.parametrize("world,gpu_mode", [(2, True), (3, False)])
def test_fake_shards_on_one_device(world, gpu_mode):
  """dynamic_embedding_ops_test.py:324-440: shards faked on one device; every key lands on owner(key)"""
  rng = np.random.default_rng(world)
  dim = 8
  tables = [Table(dim=dim, init=4096, max_capacity=4096) for _ in range(world)]
  g = PeerGroup(tables, None, world, 0, gpu_mode)
  keys = rng.integers(-2**62, 2**62, size=1500, dtype=np.int64)
  keys = np.unique(keys)
  vals = rng.standard_normal((len(keys), dim)).astype(np.float32)
  g.insert(keys, vals)
  owner = O.default_partition_fn(keys, world, gpu_mode)
  for s in range(world):
    assert tables[s].size() == int((owner == s).sum())
    out, ex = tables[s].find(keys[owner == s])
    assert ex.all()
    np.testing.assert_array_equal(out, vals[owner == s])
  q = np.concatenate([keys[::3], np.array([12345, -7], dtype=np.int64)])
  out, ex = g.find(q, dim, default=np.full(dim, 0.5, np.float32))
  assert ex[:-2].all() and not ex[-2:].any()
  np.testing.assert_array_equal(out[:-2], vals[::3])
  assert (out[-2:] == 0.5).all()
  g.close()
  for t in tables:
    t.close()


def test_multi_rank_group_threads_as_ranks():
  """3 ranks (threads), each with its own shard and peer group; peers are mapped through exported handles.
  forward: everybody inserts its batch, barrier, everybody looks up everybody's keys;
  backward: route (key, row) pairs to their owners, barrier, owners drain their inbox."""
  world, dim, per = 3, 4, 400
  rng = np.random.default_rng(5)
  all_keys = rng.choice(1 << 40, size=world * per, replace=False).astype(np.int64)
  all_vals = rng.standard_normal((world * per, dim)).astype(np.float32)
  owner = O.default_partition_fn(all_keys, world, True)
  tables = [Table(dim=dim, init=4096, max_capacity=4096) for _ in range(world)]
  hb = X().det_peer_handle_bytes()
  blob = (ctypes.c_ubyte * (hb * world))()
  for r in range(world):
    ck(X().det_peer_export(tables[r].h, ctypes.c_void_p(ctypes.addressof(blob) + r * hb)))
  rb = dim * 4
  nbytes = X().det_peer_inbox_bytes(world, 1024, rb)
  raw = [np.zeros(nbytes + 256, dtype=np.uint8) for _ in range(world)]
  boxes = [b[(-b.ctypes.data) % 256:][:nbytes] for b in raw]         # 256 B aligned views (zeroed by their owner)
  results, errors = {}, []
  start = threading.Barrier(world)

  def rank_main(r):
    try:
      tl = [None] * world
      tl[r] = tables[r]
      g = PeerGroup(tl, ctypes.cast(blob, ctypes.c_void_p), world, r)
      ptrs = (ctypes.c_void_p * world)(*[b.ctypes.data for b in boxes])
      ck(X().det_peer_inbox_attach(g.g, ptrs, 1024, rb))
      start.wait()
      mine = slice(r * per, (r + 1) * per)
      g.insert(all_keys[mine], all_vals[mine])          # writes land on their owners, over "NVLink"
      g.barrier()                                       # all ranks wrote -> all ranks may read
      out, ex = g.find(all_keys, dim)
      g.barrier()
      # backward: my batch's "gradients" travel to the owners
      keys = np.ascontiguousarray(all_keys[mine])
      rows = np.ascontiguousarray(all_vals[mine] * 2)
      ck(X().det_peer_route(g.g, P(keys), P(rows), per, None))
      g.barrier()
      counts = (ctypes.c_int64 * world)()
      ck(X().det_peer_inbox_counts(g.g, r, counts, None))
      total = int(sum(counts))
      rk = np.empty(total, dtype=np.int64)
      rr = np.empty((total, dim), dtype=np.float32)
      ck(X().det_peer_inbox_gather(g.g, r, counts, P(rk), P(rr), None))
      g.barrier()
      results[r] = (out, ex, rk, rr, list(counts))
      g.close()
    except Exception as e:  # pragma: no cover
      errors.append((r, repr(e)))
      try:
        start.abort()
      except Exception:
        pass

  th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
  for x in th:
    x.start()
  for x in th:
    x.join(timeout=300)
  assert not errors, errors
  assert len(results) == world
  for r in range(world):
    out, ex, rk, rr, counts = results[r]
    assert ex.all()
    np.testing.assert_array_equal(out, all_vals)                        # every rank sees every key
    assert tables[r].size() == int((owner == r).sum())
    exp = owner == r
    assert len(rk) == int(exp.sum())
    assert counts == [int((owner[s * per:(s + 1) * per] == r).sum()) for s in range(world)]
    o1, o2 = np.argsort(rk), np.argsort(all_keys[exp])
    np.testing.assert_array_equal(rk[o1], all_keys[exp][o2])
    np.testing.assert_array_equal(rr[o1], (all_vals[exp] * 2)[o2])
  for t in tables:
    st = t.stats()
    assert st["error_flags"] == 0
    t.close()


@pytest.mark.parametrize("world,dim", [(2, 4), (3, 16)])
def test_owner_side_exchange_threads_as_ranks(world, dim):
  """det_peer_xchg_find / det_peer_xchg_insert (the push-only sharded path): every rank a thread with its own shard,
  group and mailbox.  Several rounds of insert -> lookup with changing values, misses with a broadcast default and with
  per-key defaults, exists masks, ranks with DIFFERENT batch sizes (one of them empty), no barrier call anywhere:
  the flag words alone order owner-side reads and writes.  Every lookup must see exactly the dict model."""
  per, rounds, cap = 300, 4, 512
  rng = np.random.default_rng(11 + world)
  pool = rng.choice(1 << 40, size=world * per, replace=False).astype(np.int64)
  owner = O.default_partition_fn(pool, world, True)
  tables = [Table(dim=dim, init=4096, max_capacity=4096) for _ in range(world)]
  hb = X().det_peer_handle_bytes()
  blob = (ctypes.c_ubyte * (hb * world))()
  for r in range(world):
    ck(X().det_peer_export(tables[r].h, ctypes.c_void_p(ctypes.addressof(blob) + r * hb)))
  rb = dim * 4
  nbytes = X().det_peer_xchg_bytes(world, cap, rb)
  assert nbytes > 0 and X().det_peer_xchg_bytes(9, cap, rb) == 0
  raw = [np.zeros(nbytes + 256, dtype=np.uint8) for _ in range(world)]
  boxes = [b[(-b.ctypes.data) % 256:][:nbytes] for b in raw]
  # the schedule, known to every rank: round -> per rank (keys inserted, values) ; model applied in rank order is NOT
  # needed because a key is written by exactly one rank per round
  sched = []
  for t in range(rounds):
    perm = rng.permutation(world * per)
    parts = np.array_split(perm[: world * per - 50 * t], world)          # differing sizes per rank and round
    if t == 1:
      parts[0] = parts[0][:0]                                             # rank 0 sends nothing in round 1
    vals = [rng.standard_normal((len(q), dim)).astype(np.float32) for q in parts]
    sched.append((parts, vals))
  errors, results = [], {}
  start = threading.Barrier(world)

  def rank_main(r):
    try:
      tl = [None] * world
      tl[r] = tables[r]
      g = PeerGroup(tl, ctypes.cast(blob, ctypes.c_void_p), world, r)
      ptrs = (ctypes.c_void_p * world)(*[b.ctypes.data for b in boxes])
      ck(X().det_peer_xchg_attach(g.g, ptrs, cap, rb))
      assert X().det_peer_xchg_attach(g.g, ptrs, cap, rb + 4) != 0       # row size must match the shards
      start.wait()
      model, got = {}, []
      for t in range(rounds):
        parts, vals = sched[t]
        keys = np.ascontiguousarray(pool[parts[r]])
        v = np.ascontiguousarray(vals[r])
        ck(X().det_peer_xchg_insert(g.g, P(keys) if len(keys) else None, P(v) if len(keys) else None, len(keys), None))
        for q in range(world):                                           # what the whole job wrote this round
          for k, row in zip(pool[parts[q]].tolist(), vals[q]):
            model[k] = row
        # lookup: a rank-specific sample of everything + two keys nobody owns
        m = 200 + 37 * r if not (t == 2 and r == world - 1) else 0      # one rank looks up nothing in round 2
        q = np.concatenate([rng_r[r].choice(pool, size=m, replace=False), np.array([-5 - r, 1 << 50], np.int64)]) if m else np.zeros(0, np.int64)
        n = len(q)
        out = np.full((max(n, 1), dim), np.nan, np.float32)
        ex = np.zeros(max(n, 1), np.uint8)
        view = ctypes.c_void_p()
        if t % 2 == 0:
          d = np.full(dim, 0.25, np.float32)                              # broadcast default (same on every rank)
          ck(X().det_peer_xchg_find(g.g, P(q) if n else None, n, P(d), 0, P(out), P(ex), ctypes.byref(view), None))
          exp_miss = lambda j: d
        else:
          d = rng_r[r].standard_normal((max(n, 1), dim)).astype(np.float32)   # per-key defaults (full size)
          ck(X().det_peer_xchg_find(g.g, P(q) if n else None, n, P(d), 1, P(out), P(ex), ctypes.byref(view), None))
          exp_miss = lambda j: d[j]
        for j in range(n):
          k = int(q[j])
          if k in model:
            assert ex[j] == 1
            np.testing.assert_array_equal(out[j], model[k])
          else:
            assert ex[j] == 0
            np.testing.assert_array_equal(out[j], exp_miss(j))
        if n:   # the zero-copy ring view holds the same rows
          ring = np.ctypeslib.as_array(ctypes.cast(view.value, ctypes.POINTER(ctypes.c_float)), shape=(n, dim))
          np.testing.assert_array_equal(ring, out[:n])
        got.append(n)
      results[r] = got
      g.close()
    except BaseException as e:  # pragma: no cover
      import traceback
      errors.append((r, traceback.format_exc()))
      try:
        start.abort()
      except Exception:
        pass

  rng_r = [np.random.default_rng(100 + r) for r in range(world)]
  th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
  for x in th:
    x.start()
  for x in th:
    x.join(timeout=600)
  assert not errors, errors
  assert len(results) == world
  written = set()
  for parts, _ in sched:
    for q in parts:
      written.update(pool[q].tolist())
  for r in range(world):
    assert tables[r].size() == sum(1 for k in written if O.default_partition_fn(np.array([k]), world, True)[0] == r)
    assert tables[r].stats()["error_flags"] == 0
    tables[r].close()


@pytest.mark.parametrize("opt", ["adagrad", "adam"])
def test_owner_side_sharded_optimizer_step(opt):
  """det_peer_xchg_apply_*: the backward of the sharded lookup.  3 ranks (threads); every rank sends gradients for ids of
  ALL owners, the SAME ids from several ranks (summed on the owner in source-rank order), fresh ids every step; counts
  never leave the "device".  Twin per owner: oracle tables stepped with unique -> segment_reduce -> sparse_*_step over
  the pairs in (source rank, sender order); params and slot planes of every shard bit-exact after 3 steps."""
  world, dim, per, cap, steps = 3, 16, 260, 512, 3
  planes = 1 if opt == "adagrad" else 2
  rng = np.random.default_rng(77 + planes)
  pool = rng.choice(1 << 40, size=900, replace=False).astype(np.int64)
  owner_of = lambda k: O.default_partition_fn(k, world, True)
  tables = [Table(dim=dim, init=4096, max_capacity=4096, slot_planes=planes) for _ in range(world)]
  hb = X().det_peer_handle_bytes()
  blob = (ctypes.c_ubyte * (hb * world))()
  for r in range(world):
    ck(X().det_peer_export(tables[r].h, ctypes.c_void_p(ctypes.addressof(blob) + r * hb)))
  rb = dim * 4
  nbytes = X().det_peer_xchg_bytes(world, cap, rb)
  raw = [np.zeros(nbytes + 256, dtype=np.uint8) for _ in range(world)]
  boxes = [b[(-b.ctypes.data) % 256:][:nbytes] for b in raw]
  sched = [[(np.ascontiguousarray(rng.choice(pool, size=per - 40 * r, replace=False)),
             rng.normal(0, 1e-2, (per - 40 * r, dim)).astype(np.float32)) for r in range(world)] for _ in range(steps)]
  ip, ia = np.full(dim, 0.05, np.float32), np.full(dim, 0.1, np.float32)
  errors = []
  start = threading.Barrier(world)

  def rank_main(r):
    try:
      tl = [None] * world
      tl[r] = tables[r]
      g = PeerGroup(tl, ctypes.cast(blob, ctypes.c_void_p), world, r)
      ptrs = (ctypes.c_void_p * world)(*[b.ctypes.data for b in boxes])
      ck(X().det_peer_xchg_attach(g.g, ptrs, cap, rb))
      wsb = X().det_peer_xchg_apply_workspace_bytes(g.g)
      assert wsb > 0
      wraw = np.zeros(wsb + 256, np.uint8)
      ws = wraw[(-wraw.ctypes.data) % 256:][:wsb]
      start.wait()
      for t in range(steps):
        k, gr = sched[t][r]
        if opt == "adagrad":
          ck(X().det_peer_xchg_apply_adagrad(g.g, P(k), P(gr), len(k), 0.1, 0.0, P(ip), 0.1, P(ws), wsb, None))
        else:
          alpha = float(O.adam_scalars(0.01, 0.9, 0.999, t + 1))
          ck(X().det_peer_xchg_apply_adam(g.g, P(k), P(gr), len(k), alpha, 0.9, 0.999, 1e-8, P(ip), P(ws), wsb, None))
      # a lookup through the owners sees every owner's updated rows
      q = np.ascontiguousarray(pool[::5])
      out = np.empty((len(q), dim), np.float32)
      ck(X().det_peer_xchg_find(g.g, P(q), len(q), P(ip), 0, P(out), None, None, None))
      got[r] = (q, out)
      g.close()
    except BaseException:  # pragma: no cover
      import traceback
      errors.append((r, traceback.format_exc()))
      try:
        start.abort()
      except Exception:
        pass

  got = {}
  th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
  for x in th:
    x.start()
  for x in th:
    x.join(timeout=600)
  assert not errors, errors
  from tests.test_fused_emu import _export_sorted, sorted_export
  model = {}
  for o in range(world):
    tabs = [O.PortTable(dim) for _ in range(1 + planes)]
    for t in range(steps):
      ks = np.concatenate([sched[t][s][0][owner_of(sched[t][s][0]) == o] for s in range(world)])
      gs = np.concatenate([sched[t][s][1][owner_of(sched[t][s][0]) == o] for s in range(world)])
      u, idx = O.unique_first_occurrence(ks)
      gsum = O.segment_reduce(gs, idx, len(u))
      if opt == "adagrad":
        O.sparse_adagrad_step(tabs[0], tabs[1], u, gsum, 0.1, ip, ia, 0.0)
      else:
        O.sparse_adam_step(tabs[0], tabs[1], tabs[2], u, gsum, O.adam_scalars(0.01, 0.9, 0.999, t + 1), 0.9, 0.999, 1e-8, ip)
    for plane, ot in enumerate(tabs):
      k, v = _export_sorted(tables[o], plane)
      ek, ev = sorted_export(ot)
      np.testing.assert_array_equal(k, ek)
      np.testing.assert_array_equal(v, ev)
    ek, ev = tabs[0].export()
    model.update(zip(ek.tolist(), ev))
  for r in range(world):
    q, out = got[r]
    exp = np.stack([model.get(int(k), ip) for k in q])
    np.testing.assert_array_equal(out, exp)
  for t in tables:
    assert t.stats()["error_flags"] == 0
    t.close()


def test_bounded_table_reserve_clear_import_and_load(tmp_path):
  """the remaining entry points on a table with an eviction strategy: det_reserve carries the scores, det_clear /
  det_import reset them, det_load / det_insert_host go chunk by chunk through the evicting insert"""
  rng = np.random.default_rng(2)
  dim = 64
  t = Table(dim=dim, init=256, max_capacity=1 << 13, strategy=1)                      # LFU
  keys = np.arange(100, dtype=np.int64)
  t.insert(keys, np.ones((100, dim), dtype=np.float32))
  t.insert(keys[:50], np.ones((50, dim), dtype=np.float32))
  ck(L().det_reserve(t.h, 4000, None))
  assert t.stats()["rehash_count"] >= 1
  sc = t.scores_of(keys)
  assert (sc[:50] == 2).all() and (sc[50:] == 1).all()
  t.clear()
  assert t.size() == 0 and (t.scores_of(keys) == 0).all()
  t.insert(keys, np.ones((100, dim), dtype=np.float32))
  assert (t.scores_of(keys) == 1).all()
  k2 = np.arange(1000, 1040, dtype=np.int64)
  v2 = np.full((40, dim), 3, dtype=np.float32)     # (kept alive: P() holds no reference to the array it points into)
  ck(L().det_import(t.h, P(k2), P(v2), 40, None))
  assert t.size() == 40 and (t.scores_of(k2) == 1).all() and (t.scores_of(keys) == 0).all()
  # more keys than the table may hold, from host buffers / from files: the bound holds, the newest rows are right
  big = rng.choice(1 << 40, size=9000, replace=False).astype(np.int64)
  bv = rng.standard_normal((9000, dim)).astype(np.float32)
  ck(X().det_insert_host(t.h, P(big), P(bv), 9000))
  assert t.size() <= 1 << 13 and t.stats()["evict_events"] >= 1
  out, ex = t.find(big[-1000:])
  assert ex.all()
  np.testing.assert_array_equal(out, bv[-1000:])
  prefix = str(tmp_path / "bounded")
  big.astype("<i8").tofile(prefix + "-keys")
  bv.tofile(prefix + "-values")
  ck(X().det_load(t.h, prefix.encode(), 3000, 1))
  assert 0 < t.size() <= 1 << 13
  t.check()
  t.close()


def test_one_sided_entry_points_refuse_shards_with_an_eviction_strategy():
  """a remote claim would bypass the owner's score plane and its eviction at max_capacity, a remote probe could run into
  an eviction event: such shards are served by their owners only (the next test); the one-sided calls fail loudly."""
  t = Table(dim=4, init=1024, max_capacity=1024, strategy=0)        # LRU
  g = PeerGroup([t], None, 1, 0)
  k, v = np.arange(4, dtype=np.int64), np.zeros((4, 4), np.float32)
  out, ex = np.empty((4, 4), np.float32), np.empty(4, np.uint8)
  assert X().det_peer_find(g.g, P(k), 4, P(v[0]), 0, P(out), P(ex), None) == 1      # DET_INVALID_ARGUMENT
  assert b"eviction strategy" in L().det_last_error()
  assert X().det_peer_insert(g.g, P(k), P(v), 4, None) == 1
  rb = 16
  nbytes = X().det_peer_xchg_bytes(1, 64, rb)
  raw = np.zeros(nbytes + 256, dtype=np.uint8)
  box = raw[(-raw.ctypes.data) % 256:][:nbytes]
  ck(X().det_peer_xchg_attach(g.g, (ctypes.c_void_p * 1)(box.ctypes.data), 64, rb))
  ck(X().det_peer_xchg_insert(g.g, P(k), P(v), 4, None))           # through the owner: compact -> the owner's scored insert
  assert t.size() == 4
  g.close()
  t.close()
  t = Table(dim=4, init=1024, max_capacity=1024, strategy=4, gen_scores_fn=lambda ks: ks)     # CUSTOMIZED needs scores
  g = PeerGroup([t], None, 1, 0)
  raw2 = np.zeros(nbytes + 256, dtype=np.uint8)
  box2 = raw2[(-raw2.ctypes.data) % 256:][:nbytes]
  ck(X().det_peer_xchg_attach(g.g, (ctypes.c_void_p * 1)(box2.ctypes.data), 64, rb))
  assert X().det_peer_xchg_insert(g.g, P(k), P(v), 4, None) == 5                   # DET_UNIMPLEMENTED
  assert b"CUSTOMIZED" in L().det_last_error()
  g.close()
  t.close()


@pytest.mark.parametrize("strategy", [1, 0])          # LFU, LRU
def test_owner_side_upsert_on_shards_that_evict(strategy):
  """det_peer_xchg_insert on shards with an eviction strategy (3 ranks as threads, 1024 slots each): every call every
  rank upserts the same 24 HOT keys (identical rows) plus fresh cold keys, ~3x more distinct keys than fit.  Shards stay
  under the hard bound and evict; hot keys survive with their rows; every surviving cold key has its row; no duplicates,
  no error flags; an empty batch on one rank is fine."""
  world, dim, cap, steps, slots = 3, 8, 64, 70, 1024
  rng = np.random.default_rng(900 + strategy)
  owner_of = lambda k: O.default_partition_fn(k, world, True)
  row_of = lambda k: np.repeat((k % 100003).astype(np.float32)[:, None], dim, axis=1)
  hot = rng.choice(1 << 40, size=24, replace=False).astype(np.int64)
  cold_pool = (rng.choice(1 << 40, size=world * steps * (cap - 24), replace=False).astype(np.int64) | (1 << 41))
  tables = [Table(dim=dim, init=slots, max_capacity=slots, strategy=strategy) for _ in range(world)]
  hb = X().det_peer_handle_bytes()
  blob = (ctypes.c_ubyte * (hb * world))()
  for r in range(world):
    ck(X().det_peer_export(tables[r].h, ctypes.c_void_p(ctypes.addressof(blob) + r * hb)))
  rb = dim * 4
  nbytes = X().det_peer_xchg_bytes(world, cap, rb)
  raw = [np.zeros(nbytes + 256, dtype=np.uint8) for _ in range(world)]
  boxes = [b[(-b.ctypes.data) % 256:][:nbytes] for b in raw]
  sched, c = [], 0
  for t in range(steps):
    per = []
    for r in range(world):
      n_cold = cap - 24
      k = np.ascontiguousarray(np.concatenate([hot, cold_pool[c:c + n_cold]]))
      c += n_cold
      if t % 9 == r:
        k = np.zeros(0, np.int64)
      per.append((k, row_of(k)))
    sched.append(per)
  errors, got = [], {}
  start = threading.Barrier(world)
  zero = np.zeros(dim, np.float32)

  def rank_main(r):
    try:
      tl = [None] * world
      tl[r] = tables[r]
      g = PeerGroup(tl, ctypes.cast(blob, ctypes.c_void_p), world, r)
      ptrs = (ctypes.c_void_p * world)(*[b.ctypes.data for b in boxes])
      ck(X().det_peer_xchg_attach(g.g, ptrs, cap, rb))
      start.wait()
      for t in range(steps):
        k, v = sched[t][r]
        n = len(k)
        ck(X().det_peer_xchg_insert(g.g, P(k) if n else None, P(v) if n else None, n, None))
      out = np.empty((len(hot), dim), np.float32)
      ex = np.empty(len(hot), np.uint8)
      ck(X().det_peer_xchg_find(g.g, P(hot), len(hot), P(zero), 0, P(out), P(ex), None, None))
      got[r] = (out, ex.astype(bool))
      g.close()
    except BaseException:  # pragma: no cover
      import traceback
      errors.append((r, traceback.format_exc()))
      try:
        start.abort()
      except Exception:
        pass

  th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
  for x in th:
    x.start()
  for x in th:
    x.join(timeout=1200)
  assert not errors, errors
  live = 0
  for o in range(world):
    st = tables[o].stats()
    assert st["error_flags"] == 0 and st["evict_events"] > 0, st
    tables[o].check()
    ks, vs = tables[o].export()
    live += len(ks)
    assert len(ks) <= int(slots * 0.95) and (owner_of(ks) == o).all()
    assert np.isin(hot[owner_of(hot) == o], ks).all()
    np.testing.assert_array_equal(vs, row_of(ks))
  assert c - live > 1000
  for r in range(world):
    out, ex = got[r]
    assert ex.all()
    np.testing.assert_array_equal(out, row_of(hot))


class RegionTable(Table):
  """a table whose planes live in a caller-provided region (det_table_create_in_region): what PeerShardedVariable.create
  builds inside a symmetric-memory allocation"""

  def __init__(self, dim, slots, slot_planes=0, strategy=None):
    cfg = real.DetConfig()
    cfg.value_dtype = real.DTYPE_CODES["float32"]
    cfg.dim, cfg.device, cfg.num_slot_planes = dim, 0, slot_planes
    cfg.init_capacity, cfg.max_capacity, cfg.max_load_factor = slots, slots, 0.0
    cfg.flags = 0 if strategy is None else real.flags_evict(strategy)
    nbytes = X().det_table_region_bytes(ctypes.byref(cfg))
    self._raw = np.zeros(nbytes + 256, dtype=np.uint8)
    self.region = self._raw[(-self._raw.ctypes.data) % 256:][:nbytes]
    self.h = ctypes.c_void_p()
    ck(X().det_table_create_in_region(ctypes.byref(self.h), ctypes.byref(cfg), self.region.ctypes.data, nbytes))
    self.dim, self.dtype, self.strategy = dim, np.dtype(np.float32), strategy
    self.step_per_epoch, self.gen_scores_fn, self.curr_epoch, self.curr_step = 0, None, 0, 1


def test_table_in_a_region_evicts_in_place():
  """eviction needs no second set of planes, so a fixed-capacity table in a caller-provided region may have a strategy"""
  rng = np.random.default_rng(9)
  t = RegionTable(dim=4, slots=1024, strategy=1)                 # LFU
  keys = rng.choice(1 << 40, size=3000, replace=False).astype(np.int64)
  hot = keys[:100]
  for b in range(0, 3000, 150):
    k = np.concatenate([hot, keys[b:b + 150]]) if b else keys[:150]
    k = np.unique(k)
    t.insert(k, np.repeat(k.astype(np.float32)[:, None] % 1000, 4, axis=1))
  st = t.stats()
  assert st["evict_events"] > 0 and st["error_flags"] == 0 and st["capacity"] == 1024
  t.check()
  ks, vs = t.export()
  assert np.isin(hot, ks).all() and len(ks) <= int(1024 * 0.95)
  np.testing.assert_array_equal(vs[:, 0], ks.astype(np.float32) % 1000)
  t.close()


@pytest.mark.parametrize("strategy,backing", [(1, "ipc"), (0, "ipc"), (1, "regions")])          # LFU, LRU
def test_owner_side_training_on_shards_that_evict(strategy, backing):
  """Sharded table WITH eviction on the owner-side exchange: 3 ranks (threads), every shard a fixed 1024 slots with an
  eviction strategy; 80 training steps (det_peer_xchg_apply_adagrad) send 30 HOT ids from every rank every step plus
  fresh cold ids -- ~3x more distinct ids than the shards hold.  Every shard stays under its hard bound, evicts (events
  counted), keeps no duplicate and no error flag; the hot ids are never evicted (highest count / most recent) and their
  params and accumulators equal a sequential model bit for bit; every cold id that is still there carries exactly its
  one step from the initial row."""
  world, dim, cap, steps, slots = 3, 16, 64, 80, 1024
  rng = np.random.default_rng(400 + strategy)
  owner_of = lambda k: O.default_partition_fn(k, world, True)
  hot = rng.choice(1 << 40, size=30, replace=False).astype(np.int64)
  cold_pool = (rng.choice(1 << 40, size=world * steps * (cap - 30), replace=False).astype(np.int64) | (1 << 41))
  if backing == "regions":
    tables = [RegionTable(dim, slots, slot_planes=1, strategy=strategy) for _ in range(world)]
  else:
    tables = [Table(dim=dim, init=slots, max_capacity=slots, slot_planes=1, strategy=strategy) for _ in range(world)]
    hb = X().det_peer_handle_bytes()
    blob = (ctypes.c_ubyte * (hb * world))()
    for r in range(world):
      ck(X().det_peer_export(tables[r].h, ctypes.c_void_p(ctypes.addressof(blob) + r * hb)))
  rb = dim * 4
  nbytes = X().det_peer_xchg_bytes(world, cap, rb)
  raw = [np.zeros(nbytes + 256, dtype=np.uint8) for _ in range(world)]
  boxes = [b[(-b.ctypes.data) % 256:][:nbytes] for b in raw]
  sched, c = [], 0
  for t in range(steps):
    per = []
    for r in range(world):
      n_cold = int(rng.integers(0, cap - 30 + 1))
      k = np.ascontiguousarray(np.concatenate([hot, cold_pool[c:c + n_cold]]))
      c += n_cold
      if t % 7 == r:                                  # now and then a rank has nothing to send
        k = np.zeros(0, np.int64)
      per.append((k, rng.normal(0, 1e-2, (len(k), dim)).astype(np.float32)))
    sched.append(per)
  ip = np.full(dim, 0.05, np.float32)
  errors, got = [], {}
  start = threading.Barrier(world)

  def rank_main(r):
    try:
      if backing == "regions":
        g = PeerGroup.__new__(PeerGroup)
        g.g, g.world, g.rank = ctypes.c_void_p(), world, r
        rp = (ctypes.c_void_p * world)(*[t.region.ctypes.data for t in tables])
        ck(X().det_peer_group_create_regions(ctypes.byref(g.g), tables[r].h, rp, world, r, 1))
      else:
        tl = [None] * world
        tl[r] = tables[r]
        g = PeerGroup(tl, ctypes.cast(blob, ctypes.c_void_p), world, r)
      ptrs = (ctypes.c_void_p * world)(*[b.ctypes.data for b in boxes])
      ck(X().det_peer_xchg_attach(g.g, ptrs, cap, rb))
      wsb = X().det_peer_xchg_apply_workspace_bytes(g.g)
      wraw = np.zeros(wsb + 256, np.uint8)
      ws = wraw[(-wraw.ctypes.data) % 256:][:wsb]
      start.wait()
      for t in range(steps):
        k, gr = sched[t][r]
        n = len(k)
        ck(X().det_peer_xchg_apply_adagrad(g.g, P(k) if n else None, P(gr) if n else None, n, 0.1, 0.0, P(ip), 0.1, P(ws), wsb, None))
      q = np.ascontiguousarray(np.concatenate([hot, cold_pool[:c:11]]))
      out = np.empty((len(q), dim), np.float32)
      ex = np.empty(len(q), np.uint8)
      for b in range(0, len(q), cap):                 # the same number of collective calls on every rank
        qb, ob, eb = q[b:b + cap], out[b:b + cap], ex[b:b + cap]
        ck(X().det_peer_xchg_find(g.g, P(qb), len(qb), P(ip), 0, P(ob), P(eb), None, None))
      got[r] = (q, out, ex.astype(bool))
      g.close()
    except BaseException:  # pragma: no cover
      import traceback
      errors.append((r, traceback.format_exc()))
      try:
        start.abort()
      except Exception:
        pass

  th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
  for x in th:
    x.start()
  for x in th:
    x.join(timeout=1200)
  assert not errors, errors
  f32 = np.float32
  par, acc, cold_row = {}, {}, {}
  for t in range(steps):
    gsum = {}
    for r in range(world):                          # summed on the owner in source-rank order
      for kk, row in zip(sched[t][r][0].tolist(), sched[t][r][1]):
        gsum[kk] = row.copy() if kk not in gsum else (gsum[kk] + row).astype(f32)
    for kk, gg in gsum.items():
      a1 = (acc.get(kk, np.full(dim, 0.1, f32)) + gg * gg).astype(f32)
      par[kk] = (par.get(kk, ip) - (f32(0.1) * gg) / np.sqrt(a1)).astype(f32)
      acc[kk] = a1
  hot_set = set(hot.tolist())
  total_seen = len(par)
  total_live = 0
  for o in range(world):
    st = tables[o].stats()
    assert st["error_flags"] == 0 and st["evict_events"] > 0, st
    tables[o].check()
    ks, vs = tables[o].export()
    assert len(ks) <= int(slots * 0.95)
    total_live += len(ks)
    assert (owner_of(ks) == o).all()
    mine_hot = hot[owner_of(hot) == o]
    assert np.isin(mine_hot, ks).all()              # never evicted
    for kk, row in zip(ks.tolist(), vs):
      np.testing.assert_array_equal(row, par[kk])   # hot: the whole history; cold: its single step
  assert total_seen - total_live > 1000              # ~1200 ids a shard saw beyond its soft limit are gone
  for r in range(world):
    q, out, ex = got[r]
    assert ex[:len(hot)].all() and not ex.all()
    for j, kk in enumerate(q.tolist()):
      np.testing.assert_array_equal(out[j], par[kk] if ex[j] else ip)
