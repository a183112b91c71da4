// evict_emu.cc -- runs the REAL device code of recommenders_addons_b200/csrc/{common.cuh,evict_kernels.cuh} on the
// CPU under the SIMT emulator (cuda_emu.h).  Test infrastructure only: built by tests/emu/build_emu.py into
// tests/emu/_build/libevict_emu.so and driven through ctypes by tests/test_evict_emu.py, which checks the result
// against the independent Python restatement of the layout (tests/layout_model.py, tests/evict_model.py).
// The launch sequences below mirror the host code of csrc/evict.cu (evict_lowest, evict_insert) one to one.
#include "cuda_emu.h"

#include "../../recommenders_addons_b200/csrc/evict_kernels.cuh"

using namespace det;

namespace {

struct EmuTable {
  TableView v{};
  DevState st{};
  std::vector<long long> keys;
  std::vector<unsigned char> planes[kMaxPlanes];
  std::vector<unsigned long long> sc;
  EvictDev dev{};
  int strategy = 0;
  unsigned long long epoch = 0;
  int n_slot_planes = 0;
};

RowGeom geom(unsigned row_bytes, int vec) {
  RowGeom g;
  g.row_bytes = row_bytes;
  g.vpr = row_bytes / (unsigned)vec;
  unsigned lpr = 1, sh = 0;
  while (lpr < g.vpr && lpr < 32u) {
    lpr <<= 1;
    ++sh;
  }
  g.lpr = lpr;
  g.lpr_shift = sh;
  return g;
}

ScoreRule rule_of(const EmuTable* t) {
  ScoreRule r;
  r.strategy = t->strategy;
  r.epoch = t->epoch;
  return r;
}

unsigned long long lowmask(int bits) { return bits >= 64 ? ~0ull : ((1ull << bits) - 1ull); }

__global__ void find_slots_kernel(TableView t, const long long* keys, size_t n, long long* slots_out) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp0 * 32; base < n; base += nwarps * 32) {
    const size_t i = base + lane;
    const bool valid = i < n;
    const long long key = valid ? keys[i] : 0;
    const long long slot = warp_find_slots<true>(t, key, valid, lane);
    if (valid) slots_out[i] = slot;
  }
}

}  // namespace

extern "C" {

void* emu_create(unsigned long long nb, int dim, int n_slot_planes, int strategy) {
  EmuTable* t = new EmuTable();
  const size_t cap = nb * kBucket;
  t->keys.assign(cap, kEmptyKey);
  t->sc.assign(cap + 2, 0ull);
  t->planes[0].assign((cap + 2) * (size_t)dim * 4u, 0);
  for (int p = 1; p <= n_slot_planes; ++p) t->planes[p].assign((cap + 2) * (size_t)dim * 4u, 0);
  t->v.keys = t->keys.data();
  for (int p = 0; p < kMaxPlanes; ++p) t->v.planes[p] = t->planes[p].empty() ? nullptr : t->planes[p].data();
  t->v.nb = nb;
  t->v.row_bytes = (unsigned)dim * 4u;
  t->v.dim = (unsigned)dim;
  t->v.st = &t->st;
  memset(&t->st, 0, sizeof(DevState));
  t->strategy = strategy;
  t->n_slot_planes = n_slot_planes;
  return t;
}

void emu_destroy(void* h) { delete (EmuTable*)h; }

void emu_set_epoch(void* h, unsigned long long epoch) { ((EmuTable*)h)->epoch = epoch; }

void emu_state(void* h, unsigned long long* size, unsigned long long* used, unsigned* error) {
  EmuTable* t = (EmuTable*)h;
  *size = t->st.size;
  *used = t->st.used;
  *error = t->st.error;
}

// raw planes, for the Python checker
void emu_dump(void* h, long long* keys_out, unsigned long long* scores_out, float* rows_out) {
  EmuTable* t = (EmuTable*)h;
  const size_t cap = t->v.capacity();
  memcpy(keys_out, t->keys.data(), cap * 8);
  memcpy(scores_out, t->sc.data(), (cap + 2) * 8);
  if (rows_out) memcpy(rows_out, t->planes[0].data(), (cap + 2) * (size_t)t->v.row_bytes);
}

void emu_slot_plane(void* h, int plane, float* rows_out) {
  EmuTable* t = (EmuTable*)h;
  memcpy(rows_out, t->planes[plane].data(), (t->v.capacity() + 2) * (size_t)t->v.dim * 4u);
}

// evict_insert's launch (csrc/evict.cu)
void emu_insert_scored(void* h, const long long* keys, const float* values, const unsigned long long* scores,
                       const unsigned char* may_claim, size_t n, int grid) {
  EmuTable* t = (EmuTable*)h;
  const int vec = (t->v.row_bytes % 16u == 0) ? 16 : 4;
  const RowGeom g = geom(t->v.row_bytes, vec);
  const TableView v = t->v;
  const ScoreRule rule = rule_of(t);
  unsigned long long* sc = t->sc.data();
  const int np = t->n_slot_planes;
  const unsigned char* vals = (const unsigned char*)values;
  if (vec == 16)
    emu::launch(grid, kThreadsE, [&] { insert_scored_kernel<16>(v, keys, vals, scores, may_claim, n, g, np, sc, rule); });
  else
    emu::launch(grid, kThreadsE, [&] { insert_scored_kernel<4>(v, keys, vals, scores, may_claim, n, g, np, sc, rule); });
}

void emu_touch(void* h, const long long* keys, const unsigned long long* scores, size_t n, int grid) {
  EmuTable* t = (EmuTable*)h;
  const TableView v = t->v;
  const ScoreRule rule = rule_of(t);
  unsigned long long* sc = t->sc.data();
  emu::launch(grid, kThreadsE, [&] { touch_kernel(v, keys, scores, n, sc, rule); });
}

void emu_scores_of(void* h, const long long* keys, size_t n, unsigned long long* out, int mode, int grid) {
  EmuTable* t = (EmuTable*)h;
  const TableView v = t->v;
  unsigned long long* sc = t->sc.data();
  emu::launch(grid, kThreadsE, [&] { scores_of_keys_kernel(v, keys, n, sc, out, mode); });
}

void emu_find(void* h, const long long* keys, size_t n, long long* slots_out, int grid) {
  EmuTable* t = (EmuTable*)h;
  const TableView v = t->v;
  emu::launch(grid, kThreadsE, [&] { find_slots_kernel(v, keys, n, slots_out); });
}

// evict_room's classification (csrc/evict.cu)
void emu_classify(void* h, const long long* keys, const unsigned long long* scores, size_t n, int admission,
                  unsigned char* mask_out, unsigned long long* n_new, unsigned long long* n_adm, int grid) {
  EmuTable* t = (EmuTable*)h;
  const TableView v = t->v;
  const ScoreRule rule = rule_of(t);
  unsigned long long* sc = t->sc.data();
  EvictDev* d = &t->dev;
  emu::launch(1, 256, [&] { evict_reset_kernel(d); });
  if (admission) emu::launch(grid, kThreadsE, [&] { minmax_kernel(v, sc, d); });
  emu::launch(grid, kThreadsE, [&] { classify_kernel(v, keys, scores, n, rule, admission, mask_out, d); });
  *n_new = d->n_new;
  *n_adm = d->n_adm;
}

// evict_lowest (csrc/evict.cu), launch for launch
unsigned long long emu_evict_lowest(void* h, unsigned long long k, int grid, int* rounds_out,
                                    unsigned long long* tau_out, unsigned long long* quota_out) {
  EmuTable* t = (EmuTable*)h;
  const TableView v = t->v;
  unsigned long long* sc = t->sc.data();
  EvictDev* d = &t->dev;
  if (rounds_out) *rounds_out = 0;
  emu::launch(1, 256, [&] { evict_reset_kernel(d); });
  emu::launch(grid, kThreadsE, [&] { minmax_kernel(v, sc, d); });
  const unsigned long long n_live = d->n_live, smin = d->smin, smax = d->smax;
  if (n_live == 0 || k == 0) return 0;
  if (k > n_live) k = n_live;
  int sig = 0;
  for (unsigned long long diff = smin ^ smax; diff; diff >>= 1) ++sig;
  {
    const unsigned long long prefix = smin & ~lowmask(sig);
    emu::launch(1, 32, [&] {
      if (threadIdx.x == 0) select_init_kernel(d, prefix, k);
    });
  }
  for (int hi = sig; hi > 0;) {
    const int bits = hi < kHistBits ? hi : kHistBits;
    emu::launch(grid, kThreadsE, [&] { hist_kernel(v, sc, d, hi, bits); });
    emu::launch(1, 32, [&] { pick_kernel(d, hi, bits); });
    hi -= bits;
  }
  if (tau_out) *tau_out = d->prefix;
  if (quota_out) *quota_out = d->remaining;
  emu::launch(grid, kThreadsE, [&] { evict_apply_kernel(v, sc, d); });
  const int vec = (t->v.row_bytes % 16u == 0) ? 16 : 4;
  const RowGeom g = geom(t->v.row_bytes, vec);
  const RowGeom gs = geom(t->v.dim * 4u, 4);
  const int np = t->n_slot_planes;
  for (int round = 0; round < 256; ++round) {
    d->n_moved = 0;
    d->n_erased = 0;
    if (vec == 16)
      emu::launch(grid, kThreadsE, [&] { repair_move_kernel<16>(v, sc, g, gs, np, d); });
    else
      emu::launch(grid, kThreadsE, [&] { repair_move_kernel<4>(v, sc, g, gs, np, d); });
    emu::launch(grid, kThreadsE, [&] { repair_sweep_kernel(v, sc, d); });
    if (d->n_moved == 0 && d->n_erased == 0) break;
    if (rounds_out) ++*rounds_out;
  }
  return d->n_evicted;
}

// growth: rehash into a table of new_nb buckets with the ordinary find-or-claim, then carry the scores
// (rehash_kernel lives in table.cu; its claim + row move is restated with the same primitives)
void emu_carry_scores(void* h_old, void* h_new, int grid) {
  EmuTable* o = (EmuTable*)h_old;
  EmuTable* n = (EmuTable*)h_new;
  const TableView ov = o->v, nv = n->v;
  const unsigned long long* osc = o->sc.data();
  unsigned long long* nsc = n->sc.data();
  emu::launch(grid, kThreadsE, [&] { carry_scores_kernel(ov, osc, nv, nsc); });
}

}  // extern "C"

// det_remove's effect on given slots (remove_kernel lives in table.cu): a slot whose bucket still has an EMPTY slot
// goes back to EMPTY, else it becomes a tombstone; its score is zeroed first (evict_before_remove)
extern "C" void emu_remove_slots(void* h, const long long* slots, size_t n) {
  EmuTable* t = (EmuTable*)h;
  const long long cap = (long long)t->v.capacity();
  for (size_t i = 0; i < n; ++i) {
    const long long s = slots[i];
    if (s < 0 || s >= cap) continue;
    const long long b0 = s & ~(long long)(kBucket - 1);
    bool has_empty = false;
    for (int q = 0; q < kBucket; ++q) has_empty |= t->keys[b0 + q] == kEmptyKey;
    t->keys[s] = has_empty ? kEmptyKey : kTombKey;
    t->sc[s] = 0;
    t->st.size -= 1;
    if (has_empty) t->st.used -= 1;
  }
}
