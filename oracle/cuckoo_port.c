/* oracle/cuckoo_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded restatement of the reference's CPU table path for
 * int64 keys / float32 rows.  It needs nothing from /root/reference at build or
 * run time, so it travels to the GPU box and is the parity checker there.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Parity pinning: tests/test_oracle.py drives this port and the reference's own
 * libcuckoo (oracle/_ref, built from /root/reference by oracle/Makefile) with the
 * same operation streams and requires identical results INCLUDING export order
 * (both single-threaded), and checks both against the known-answer tests of
 * dynamic_embedding_variable_test.py:394-563 (tests/golden/).
 *
 * Algorithm followed (paths under
 * /root/reference/tensorflow_recommenders_addons/dynamic_embedding/core/):
 *   hash            HybridHash<int64> = murmur3 fmix64     kernels/lookup_impl/lookup_table_op_cpu.h:90-101
 *   partial key     8-bit xor fold of the hash              lib/cuckoo/cuckoohash_map.hh:875-884
 *   buckets         i1 = h & mask; i2 = (i1 ^ (tag+1)*0xc6a4a7935bd1e995) & mask   :889-903
 *   4 slots/bucket                                          lib/cuckoo/cuckoohash_config.hh:10
 *   find            scan i1 then i2, key compare            cuckoohash_map.hh:1249-1290
 *   insert          dup check + LAST empty slot of i1, then i2   :1336-1349, 1410-1430
 *   cuckoo path     BFS over <=5 levels, pathcode base 4    :1444-1760
 *   resize          double, split bucket b into b / b+2^hp  :1779-1870 (move_bucket)
 *   upsert          insert_or_assign                        :577-590, 736-740
 *   accum           4-way exists rule of accumrase_fn       :620-633, 756-765
 *   erase / clear / size / locked-table iteration order (bucket-major, slot-minor)
 *   row semantics   TableWrapperOptimized                   kernels/lookup_impl/lookup_table_op_cpu.h:148-263
 *   op semantics    Find default rule (is_full_default), Insert, Accum, Remove, Export
 *                                                           kernels/cuckoo_hashtable_op.cc:39-308
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SLOTS 4
#define MAX_BFS_PATH_LEN 5
#define MAX_CUCKOO_COUNT 682 /* 2 * ((4^5 - 1) / 3) */

typedef struct {
  size_t hp;         /* hashpower: 2^hp buckets */
  int64_t* keys;     /* [nb*4] */
  uint8_t* partial;  /* [nb*4] */
  uint8_t* occ;      /* [nb*4] */
  float* vals;       /* [nb*4*dim] */
  size_t size;
  int64_t dim;
} port_t;

typedef struct {
  uint64_t hash;
  uint8_t partial;
} hv_t;

static uint64_t fmix64(int64_t key) {
  uint64_t k = (uint64_t)key;
  k ^= k >> 33;
  k *= UINT64_C(0xff51afd7ed558ccd);
  k ^= k >> 33;
  k *= UINT64_C(0xc4ceb9fe1a85ec53);
  k ^= k >> 33;
  return k;
}

static uint8_t partial_key(uint64_t h) {
  uint32_t h32 = (uint32_t)h ^ (uint32_t)(h >> 32);
  uint16_t h16 = (uint16_t)h32 ^ (uint16_t)(h32 >> 16);
  return (uint8_t)((uint8_t)h16 ^ (uint8_t)(h16 >> 8));
}

static hv_t hashed_key(int64_t key) {
  hv_t r;
  r.hash = fmix64(key);
  r.partial = partial_key(r.hash);
  return r;
}

static size_t hashmask(size_t hp) { return ((size_t)1 << hp) - 1; }
static size_t index_hash(size_t hp, uint64_t hv) { return hv & hashmask(hp); }
static size_t alt_index(size_t hp, uint8_t partial, size_t index) {
  const uint64_t tag = (uint64_t)partial + 1;
  return (index ^ (tag * UINT64_C(0xc6a4a7935bd1e995))) & hashmask(hp);
}

static size_t reserve_calc(size_t n) {
  size_t buckets = (n + SLOTS - 1) / SLOTS, blog2;
  for (blog2 = 0; ((size_t)1 << blog2) < buckets; ++blog2) {
  }
  return blog2;
}

static void alloc_planes(port_t* t, size_t hp) {
  size_t n = ((size_t)1 << hp) * SLOTS;
  t->hp = hp;
  t->keys = (int64_t*)malloc(n * sizeof(int64_t));
  t->partial = (uint8_t*)malloc(n);
  t->occ = (uint8_t*)calloc(n, 1);
  t->vals = (float*)malloc(n * (size_t)t->dim * sizeof(float));
}

static void free_planes(port_t* t) {
  free(t->keys);
  free(t->partial);
  free(t->occ);
  free(t->vals);
}

void* port_create(long long dim, size_t init_size) {
  port_t* t = (port_t*)malloc(sizeof(port_t));
  if (init_size == 0) init_size = 8192; /* cuckoo_hashtable_op.cc:199-205 */
  t->dim = dim;
  t->size = 0;
  alloc_planes(t, reserve_calc(init_size));
  return t;
}

void port_destroy(void* p) {
  port_t* t = (port_t*)p;
  free_planes(t);
  free(t);
}

static float* row(port_t* t, size_t b, size_t s) { return t->vals + (b * SLOTS + s) * (size_t)t->dim; }

/* try_read_from_bucket, cuckoohash_map.hh:1265-1278 (int64 keys are "simple": no tag compare) */
static int try_read_from_bucket(const port_t* t, size_t b, int64_t key) {
  for (int i = 0; i < SLOTS; ++i) {
    if (!t->occ[b * SLOTS + i]) continue;
    if (t->keys[b * SLOTS + i] == key) return i;
  }
  return -1;
}

/* cuckoo_find, :1249-1261 */
static int cuckoo_find(const port_t* t, int64_t key, size_t i1, size_t i2, size_t* ob, size_t* os) {
  int s = try_read_from_bucket(t, i1, key);
  if (s != -1) {
    *ob = i1;
    *os = (size_t)s;
    return 1;
  }
  s = try_read_from_bucket(t, i2, key);
  if (s != -1) {
    *ob = i2;
    *os = (size_t)s;
    return 1;
  }
  return 0;
}

/* try_find_insert_bucket, :1410-1430 : returns 0 on duplicate (slot = dup), else 1 with slot =
 * LAST empty slot or -1 */
static int try_find_insert_bucket(const port_t* t, size_t b, int* slot, int64_t key) {
  *slot = -1;
  for (int i = 0; i < SLOTS; ++i) {
    if (t->occ[b * SLOTS + i]) {
      if (t->keys[b * SLOTS + i] == key) {
        *slot = i;
        return 0;
      }
    } else {
      *slot = i;
    }
  }
  return 1;
}

typedef struct {
  size_t bucket;
  uint16_t pathcode;
  int8_t depth;
} b_slot;

typedef struct {
  size_t bucket;
  size_t slot;
  hv_t hv;
} cuckoo_record;

/* slot_search, :1727-1760 */
static b_slot slot_search(const port_t* t, size_t i1, size_t i2) {
  b_slot q[MAX_CUCKOO_COUNT];
  size_t first = 0, last = 0;
  b_slot r;
  q[last].bucket = i1, q[last].pathcode = 0, q[last].depth = 0, last++;
  q[last].bucket = i2, q[last].pathcode = 1, q[last].depth = 0, last++;
  while (first != last) {
    b_slot x = q[first++];
    size_t starting_slot = x.pathcode % SLOTS;
    for (size_t i = 0; i < SLOTS; ++i) {
      uint16_t slot = (uint16_t)((starting_slot + i) % SLOTS);
      if (!t->occ[x.bucket * SLOTS + slot]) {
        x.pathcode = (uint16_t)(x.pathcode * SLOTS + slot);
        return x;
      }
      if (x.depth < MAX_BFS_PATH_LEN - 1) {
        b_slot y;
        y.bucket = alt_index(t->hp, t->partial[x.bucket * SLOTS + slot], x.bucket);
        y.pathcode = (uint16_t)(x.pathcode * SLOTS + slot);
        y.depth = (int8_t)(x.depth + 1);
        q[last++] = y;
      }
    }
  }
  r.bucket = 0, r.pathcode = 0, r.depth = -1;
  return r;
}

/* cuckoopath_search, :1553-1611 */
static int cuckoopath_search(const port_t* t, cuckoo_record* path, size_t i1, size_t i2) {
  b_slot x = slot_search(t, i1, i2);
  if (x.depth == -1) return -1;
  for (int i = x.depth; i >= 0; i--) {
    path[i].slot = x.pathcode % SLOTS;
    x.pathcode /= SLOTS;
  }
  path[0].bucket = (x.pathcode == 0) ? i1 : i2;
  if (!t->occ[path[0].bucket * SLOTS + path[0].slot]) return 0;
  path[0].hv = hashed_key(t->keys[path[0].bucket * SLOTS + path[0].slot]);
  for (int i = 1; i <= x.depth; ++i) {
    path[i].bucket = alt_index(t->hp, path[i - 1].hv.partial, path[i - 1].bucket);
    if (!t->occ[path[i].bucket * SLOTS + path[i].slot]) return i;
    path[i].hv = hashed_key(t->keys[path[i].bucket * SLOTS + path[i].slot]);
  }
  return x.depth;
}

/* cuckoopath_move, :1620-1682 */
static int cuckoopath_move(port_t* t, cuckoo_record* path, int depth) {
  if (depth == 0) return !t->occ[path[0].bucket * SLOTS + path[0].slot];
  while (depth > 0) {
    cuckoo_record* from = &path[depth - 1];
    cuckoo_record* to = &path[depth];
    size_t fi = from->bucket * SLOTS + from->slot, ti = to->bucket * SLOTS + to->slot;
    if (t->occ[ti] || !t->occ[fi] || fmix64(t->keys[fi]) != from->hv.hash) return 0;
    t->keys[ti] = t->keys[fi];
    t->partial[ti] = t->partial[fi];
    memcpy(t->vals + ti * (size_t)t->dim, t->vals + fi * (size_t)t->dim, (size_t)t->dim * sizeof(float));
    t->occ[ti] = 1;
    t->occ[fi] = 0;
    depth--;
  }
  return 1;
}

/* run_cuckoo, :1458-1492 : 1 = freed (bucket, slot), 0 = table full */
static int run_cuckoo(port_t* t, size_t i1, size_t i2, size_t* ib, size_t* is) {
  cuckoo_record path[MAX_BFS_PATH_LEN];
  while (1) {
    int depth = cuckoopath_search(t, path, i1, i2);
    if (depth < 0) return 0;
    if (cuckoopath_move(t, path, depth)) {
      *ib = path[0].bucket;
      *is = path[0].slot;
      return 1;
    }
  }
}

/* move_bucket / cuckoo_fast_double, :1779-1870 */
static void fast_double(port_t* t) {
  port_t old = *t;
  size_t old_hp = old.hp, new_hp = old.hp + 1, nb_old = (size_t)1 << old_hp;
  alloc_planes(t, new_hp);
  for (size_t b = 0; b < nb_old; ++b) {
    size_t new_bucket_ind = b + nb_old, new_bucket_slot = 0;
    for (size_t s = 0; s < SLOTS; ++s) {
      size_t oi = b * SLOTS + s;
      if (!old.occ[oi]) continue;
      hv_t hv = hashed_key(old.keys[oi]);
      size_t old_ihash = index_hash(old_hp, hv.hash);
      size_t old_ahash = alt_index(old_hp, hv.partial, old_ihash);
      size_t new_ihash = index_hash(new_hp, hv.hash);
      size_t new_ahash = alt_index(new_hp, hv.partial, new_ihash);
      size_t db, ds;
      if ((b == old_ihash && new_ihash == new_bucket_ind) ||
          (b == old_ahash && new_ahash == new_bucket_ind)) {
        db = new_bucket_ind;
        ds = new_bucket_slot++;
      } else {
        db = b;
        ds = s;
      }
      size_t ni = db * SLOTS + ds;
      t->keys[ni] = old.keys[oi];
      t->partial[ni] = old.partial[oi];
      t->occ[ni] = 1;
      memcpy(t->vals + ni * (size_t)t->dim, old.vals + oi * (size_t)t->dim,
             (size_t)t->dim * sizeof(float));
    }
  }
  free_planes(&old);
}

/* cuckoo_insert_loop + cuckoo_insert, :1295-1400 : returns 1 = free position (new key),
 * 0 = duplicate at position */
static int insert_position(port_t* t, int64_t key, hv_t hv, size_t* ob, size_t* os) {
  while (1) {
    size_t i1 = index_hash(t->hp, hv.hash);
    size_t i2 = alt_index(t->hp, hv.partial, i1);
    int res1, res2;
    if (!try_find_insert_bucket(t, i1, &res1, key)) {
      *ob = i1, *os = (size_t)res1;
      return 0;
    }
    if (!try_find_insert_bucket(t, i2, &res2, key)) {
      *ob = i2, *os = (size_t)res2;
      return 0;
    }
    if (res1 != -1) {
      *ob = i1, *os = (size_t)res1;
      return 1;
    }
    if (res2 != -1) {
      *ob = i2, *os = (size_t)res2;
      return 1;
    }
    if (run_cuckoo(t, i1, i2, ob, os)) return 1;
    fast_double(t);
  }
}

static void add_to_bucket(port_t* t, size_t b, size_t s, hv_t hv, int64_t key, const float* v) {
  size_t i = b * SLOTS + s;
  t->keys[i] = key;
  t->partial[i] = hv.partial;
  t->occ[i] = 1;
  memcpy(row(t, b, s), v, (size_t)t->dim * sizeof(float));
  t->size++;
}

/* LaunchTensorsFind(+WithExists), cuckoo_hashtable_op.cc:39-106 over TableWrapperOptimized::find */
void port_find(void* p, const long long* keys, long long n, const float* defaults, int full_default,
               float* out, unsigned char* exists) {
  port_t* t = (port_t*)p;
  size_t dim = (size_t)t->dim;
  for (long long i = 0; i < n; ++i) {
    hv_t hv = hashed_key(keys[i]);
    size_t i1 = index_hash(t->hp, hv.hash), i2 = alt_index(t->hp, hv.partial, i1), b, s;
    int found = cuckoo_find(t, keys[i], i1, i2, &b, &s);
    if (found) {
      memcpy(out + (size_t)i * dim, row(t, b, s), dim * sizeof(float));
    } else {
      const float* d = full_default ? defaults + (size_t)i * dim : defaults;
      for (size_t j = 0; j < dim; ++j) out[(size_t)i * dim + j] = d[j];
    }
    if (exists) exists[i] = (unsigned char)found;
  }
}

/* LaunchTensorsInsert -> insert_or_assign (uprase_fn), cuckoohash_map.hh:577-590 */
void port_insert(void* p, const long long* keys, const float* values, long long n) {
  port_t* t = (port_t*)p;
  size_t dim = (size_t)t->dim;
  for (long long i = 0; i < n; ++i) {
    hv_t hv = hashed_key(keys[i]);
    size_t b, s;
    if (insert_position(t, keys[i], hv, &b, &s)) {
      add_to_bucket(t, b, s, hv, keys[i], values + (size_t)i * dim);
    } else {
      memcpy(row(t, b, s), values + (size_t)i * dim, dim * sizeof(float));
    }
  }
}

/* LaunchTensorsAccum -> insert_or_accum (accumrase_fn), cuckoohash_map.hh:620-633:
 *   (not found, !exist) -> insert row;  (found, exist) -> row += delta;  otherwise no-op.
 * NOTE the reference runs cuckoo_insert_loop before looking at `exist`, so a (not found, exist)
 * call can still displace entries / double the table; it just does not add the key. */
void port_accum(void* p, const long long* keys, const float* vod, const unsigned char* exists,
                long long n) {
  port_t* t = (port_t*)p;
  size_t dim = (size_t)t->dim;
  for (long long i = 0; i < n; ++i) {
    hv_t hv = hashed_key(keys[i]);
    size_t b, s;
    int is_new = insert_position(t, keys[i], hv, &b, &s);
    if (is_new && !exists[i]) {
      add_to_bucket(t, b, s, hv, keys[i], vod + (size_t)i * dim);
    } else if (!is_new && exists[i]) {
      float* r = row(t, b, s);
      for (size_t j = 0; j < dim; ++j) r[j] += vod[(size_t)i * dim + j]; /* ValueArray::operator+= */
    }
  }
}

/* Remove, cuckoo_hashtable_op.cc:268-276 */
void port_remove(void* p, const long long* keys, long long n) {
  port_t* t = (port_t*)p;
  for (long long i = 0; i < n; ++i) {
    hv_t hv = hashed_key(keys[i]);
    size_t i1 = index_hash(t->hp, hv.hash), i2 = alt_index(t->hp, hv.partial, i1), b, s;
    if (cuckoo_find(t, keys[i], i1, i2, &b, &s)) {
      t->occ[b * SLOTS + s] = 0;
      t->size--;
    }
  }
}

void port_clear(void* p) {
  port_t* t = (port_t*)p;
  memset(t->occ, 0, ((size_t)1 << t->hp) * SLOTS);
  t->size = 0;
}

size_t port_size(void* p) { return ((port_t*)p)->size; }

/* dump(), lookup_table_op_cpu.h:219-252 : locked-table iteration = bucket-major, slot-minor */
size_t port_export(void* p, long long* keys, float* values, size_t offset, size_t length) {
  port_t* t = (port_t*)p;
  size_t dim = (size_t)t->dim, n = ((size_t)1 << t->hp) * SLOTS, seen = 0, out = 0;
  if (offset > t->size || t->size == 0) return 0;
  for (size_t i = 0; i < n && out < length; ++i) {
    if (!t->occ[i]) continue;
    if (seen++ < offset) continue;
    keys[out] = t->keys[i];
    memcpy(values + out * dim, t->vals + i * dim, dim * sizeof(float));
    out++;
  }
  return out;
}
