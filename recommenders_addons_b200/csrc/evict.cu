// evict.cu -- capacity management of a table with a fixed max_capacity: per-key SCORES (one more plane, co-indexed
// by slot) and score-based eviction.  This is the HkvHashTable side of the reference
// (DE/core/kernels/lookup_impl/lookup_table_op_hkv.h:436-547: strategies LRU / LFU / EPOCHLRU / EPOCHLFU /
// CUSTOMIZED, scores input of Insert/Accum, export with scores; tests kernel_tests/hkv_hashtable_evict_test.py).
//
// B200-first design (DESIGN.md "capacity management"): HierarchicalKV evicts inside every insert, from the key's
// own 128-slot bucket.  Here the hot kernels stay exactly the growth-mode kernels (short probe chains at load
// <= 0.875, no per-insert score scans); eviction is an AMORTISED maintenance event that streams the score plane
// at HBM speed:
//   1. min/max + live count of the scores                          (1 pass over keys+scores)
//   2. radix select of the k-th lowest score, 11 bits per pass, only over the bits in which min and max differ
//      (LFU counts: 1 pass; all scores equal: 0 passes)            (<= 6 passes)
//   3. erase every key below the threshold (+ a quota of the ties) straight to EMPTY
//   4. repair rounds: a key whose probe chain was cut by (3) is re-seated in the first free slot of its chain
//      (closer to home), until a round moves nothing -- eviction never leaves tombstones behind
// k = what the incoming batch needs + a slab (1/32 of the limit), so that the next event is many steps away.
// A key that is not in the table is refused when its score is below every resident score (HKV refuses a key
// whose score is below its bucket's minimum): det_insert_scored gets a may-claim mask from the classify pass.
// The sequential model of this file is tests/evict_model.py (property tests + the reference's eviction tests).
#include <algorithm>
#include <vector>

#include "host.h"

namespace det {

constexpr int kThreadsE = 256;
constexpr int kHistBits = 11;
constexpr int kHistBins = 1 << kHistBits;
constexpr unsigned long long kM32 = 0xffffffffull;

struct EvictDev {
  unsigned long long smin, smax, n_live;        // pass 1
  unsigned long long prefix, remaining;         // radix select: decided high bits / rank left inside the prefix
  unsigned long long n_new, n_adm;              // classify: keys of the batch not in the table / of those, admitted
  unsigned long long tie_ticket, n_evicted, n_moved, n_erased;
  unsigned long long pad[5];
  unsigned int hist[kHistBins];
};

struct ScoreRule {
  int strategy;
  unsigned long long epoch;
};

struct EvictState {
  int strategy = 0;
  unsigned long long* scores = nullptr;  // device [capacity + 2]
  unsigned long long epoch = 0;
  EvictDev* dev = nullptr;
  EvictDev* h_dev = nullptr;  // pinned mirror
  // the mutating call being served (set by the entry point under the table mutex, read by evict_room)
  const unsigned long long* ctx_scores = nullptr;
  bool ctx_admission = false;
  unsigned char* ctx_mask = nullptr;  // out: may-claim mask of the launch (device), or null = every key may claim
  uint32_t n_events = 0;
  uint64_t n_evicted = 0;
};

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// score of a key after an insert / assign / accumulate (HierarchicalKV v0.1.0-beta.12 rules)
__device__ __forceinline__ unsigned long long rule_score(const ScoreRule& r, unsigned long long old, bool has,
                                                         unsigned long long provided, unsigned long long now) {
  switch (r.strategy) {
    case DET_EVICT_LRU: return now;
    case DET_EVICT_LFU: return old + (has ? provided : 1ull);
    case DET_EVICT_EPOCHLRU: return (r.epoch << 32) | ((now >> 20) & kM32);
    case DET_EVICT_EPOCHLFU: {
      const unsigned long long d = has ? (provided > kM32 ? kM32 : provided) : 1ull;
      unsigned long long f = (old & kM32) + d;
      if (f > kM32) f = kM32;
      return (r.epoch << 32) | f;
    }
    default: return has ? provided : old;  // CUSTOMIZED
  }
}

__device__ __forceinline__ bool live_key_e(long long k) { return k != kEmptyKey && k != kTombKey; }

// L2-coherent load / store of one key (repair rounds: other warps claim and free slots meanwhile)
__device__ __forceinline__ long long ld_key_cg(const long long* p) {
  long long r;
  asm volatile("ld.global.cg.s64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_key_cg(long long* p, long long v) {
  asm volatile("st.global.cg.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- hot path: insert_or_assign that also writes the score plane ---------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreadsE)
insert_scored_kernel(TableView t, const long long* __restrict__ keys, const unsigned char* __restrict__ values,
                     const unsigned long long* __restrict__ scores_in, const unsigned char* __restrict__ may_claim,
                     size_t n, RowGeom g, int n_slot_planes, unsigned long long* __restrict__ sc, ScoreRule rule) {
  __shared__ unsigned s_new, s_used;
  if (threadIdx.x == 0) {
    s_new = 0;
    s_used = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsE + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsE) >> 5;
  for (size_t base = warp0 * 32; base < n; base += nwarps * 32) {
    const size_t i = base + lane;
    const bool valid = i < n;
    const long long key = valid ? __ldg(keys + i) : 0;
    const bool claim = valid && (may_claim == nullptr || may_claim[i] != 0);
    bool is_new, from_empty;
    const long long slot = warp_find_or_claim(t, key, valid, claim, lane, is_new, from_empty);
    const unsigned bn = __ballot_sync(kFull, is_new), bu = __ballot_sync(kFull, from_empty);
    if (lane == 0 && bn) {
      atomicAdd(&s_new, __popc(bn));
      atomicAdd(&s_used, __popc(bu));
    }
    const unsigned char* src = nullptr;
    unsigned char* dst = nullptr;
    if (valid && slot >= 0) {
      src = values + i * g.row_bytes;
      dst = t.planes[0] + (size_t)slot * g.row_bytes;
    }
    warp_move_rows<VEC>(g, src, dst, lane);
    if (is_new && slot >= 0)
      for (int p = 1; p <= n_slot_planes; ++p)
        *reinterpret_cast<unsigned*>(t.planes[p] + (size_t)slot * t.dim * 4u) = kSlotUninit;
    if (valid && slot >= 0) {
      const unsigned long long old = is_new ? 0ull : sc[slot];
      sc[slot] = rule_score(rule, old, scores_in != nullptr, scores_in ? scores_in[i] : 0ull, now_ns());
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_new) {
    atomicAdd(&t.st->size, (unsigned long long)s_new);
    atomicAdd(&t.st->used, (unsigned long long)s_used);
  }
}

// score update of the keys a mutating kernel of table.cu / fused.cu has just written (accum, fused optimizer).
// A key created by that kernel sits in a slot whose score is 0 (free slots always carry score 0).
__global__ void __launch_bounds__(kThreadsE)
touch_kernel(TableView t, const long long* __restrict__ keys, const unsigned long long* __restrict__ scores_in,
             size_t n, unsigned long long* __restrict__ sc, ScoreRule rule) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsE + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsE) >> 5;
  for (size_t base = warp0 * 32; base < n; base += nwarps * 32) {
    const size_t i = base + lane;
    const bool valid = i < n;
    const long long key = valid ? __ldg(keys + i) : 0;
    const long long slot = warp_find_slots<true>(t, key, valid, lane);
    if (valid && slot >= 0)
      sc[slot] = rule_score(rule, sc[slot], scores_in != nullptr, scores_in ? scores_in[i] : 0ull, now_ns());
  }
}

// mode 0: scores_out[i] = score of keys[i] (0 when absent);  mode 1: zero the score of keys[i] (before a remove)
__global__ void __launch_bounds__(kThreadsE)
scores_of_keys_kernel(TableView t, const long long* __restrict__ keys, size_t n, unsigned long long* __restrict__ sc,
                      unsigned long long* __restrict__ scores_out, int mode) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsE + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsE) >> 5;
  for (size_t base = warp0 * 32; base < n; base += nwarps * 32) {
    const size_t i = base + lane;
    const bool valid = i < n;
    const long long key = valid ? __ldg(keys + i) : 0;
    const long long slot = warp_find_slots<true>(t, key, valid, lane);
    if (!valid) continue;
    if (mode == 0)
      scores_out[i] = slot >= 0 ? sc[slot] : 0ull;
    else if (slot >= 0)
      sc[slot] = 0ull;
  }
}

// ---- eviction event ------------------------------------------------------------------------------------------
__global__ void evict_reset_kernel(EvictDev* d) {
  for (int b = threadIdx.x; b < kHistBins; b += blockDim.x) d->hist[b] = 0;
  if (threadIdx.x == 0) {
    d->smin = ~0ull;
    d->smax = 0;
    d->n_live = 0;
    d->prefix = 0;
    d->remaining = 0;
    d->n_new = 0;
    d->n_adm = 0;
    d->tie_ticket = 0;
    d->n_evicted = 0;
    d->n_moved = 0;
    d->n_erased = 0;
  }
}

__global__ void __launch_bounds__(kThreadsE)
minmax_kernel(TableView t, const unsigned long long* __restrict__ sc, EvictDev* d) {
  const size_t cap = t.capacity();
  unsigned long long mn = ~0ull, mx = 0, cnt = 0;
  for (size_t s = (size_t)blockIdx.x * kThreadsE + threadIdx.x; s < cap; s += (size_t)gridDim.x * kThreadsE) {
    if (live_key_e(t.keys[s])) {
      const unsigned long long v = sc[s];
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
      ++cnt;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_down_sync(kFull, mn, o), b = __shfl_down_sync(kFull, mx, o),
                             c = __shfl_down_sync(kFull, cnt, o);
    mn = a < mn ? a : mn;
    mx = b > mx ? b : mx;
    cnt += c;
  }
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicMin(&d->smin, mn);
    atomicMax(&d->smax, mx);
    atomicAdd(&d->n_live, cnt);
  }
}

// which keys of the batch are new, and which of those may be admitted (score >= lowest resident score)
__global__ void __launch_bounds__(kThreadsE)
classify_kernel(TableView t, const long long* __restrict__ keys, const unsigned long long* __restrict__ scores_in,
                size_t n, ScoreRule rule, int admission, unsigned char* __restrict__ mask_out, EvictDev* d) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsE + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsE) >> 5;
  const unsigned long long smin = d->n_live ? d->smin : 0ull;
  unsigned c_new = 0, c_adm = 0;
  for (size_t base = warp0 * 32; base < n; base += nwarps * 32) {
    const size_t i = base + lane;
    const bool valid = i < n;
    const long long key = valid ? __ldg(keys + i) : 0;
    const long long slot = warp_find_slots<true>(t, key, valid, lane);
    const bool missing = valid && slot < 0;
    bool adm = missing;
    if (missing && admission)
      adm = rule_score(rule, 0ull, scores_in != nullptr, scores_in ? scores_in[i] : 0ull, now_ns()) >= smin;
    if (valid && mask_out) mask_out[i] = (slot >= 0 || adm) ? 1 : 0;
    c_new += __popc(__ballot_sync(kFull, missing));
    c_adm += __popc(__ballot_sync(kFull, adm));
  }
  if (lane == 0 && c_new) {
    atomicAdd(&d->n_new, (unsigned long long)c_new);
    atomicAdd(&d->n_adm, (unsigned long long)c_adm);
  }
}

__global__ void select_init_kernel(EvictDev* d, unsigned long long prefix, unsigned long long k) {
  d->prefix = prefix;
  d->remaining = k;
  d->tie_ticket = 0;
  d->n_evicted = 0;
}

// histogram of bits [hi-bits, hi) of the live scores whose bits >= hi equal the prefix decided so far
__global__ void __launch_bounds__(kThreadsE)
hist_kernel(TableView t, const unsigned long long* __restrict__ sc, EvictDev* d, int hi, int bits) {
  __shared__ unsigned h[kHistBins];
  for (int b = threadIdx.x; b < kHistBins; b += kThreadsE) h[b] = 0;
  __syncthreads();
  const size_t cap = t.capacity();
  const unsigned long long prefix = d->prefix;
  const int shift = hi - bits;
  const unsigned bmask = (1u << bits) - 1u;
  for (size_t s = (size_t)blockIdx.x * kThreadsE + threadIdx.x; s < cap; s += (size_t)gridDim.x * kThreadsE) {
    if (!live_key_e(t.keys[s])) continue;
    const unsigned long long v = sc[s];
    if (hi < 64 && (v >> hi) != (prefix >> hi)) continue;
    atomicAdd(&h[(unsigned)(v >> shift) & bmask], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kHistBins; b += kThreadsE)
    if (h[b]) atomicAdd(&d->hist[b], h[b]);
}

// the bin that holds the `remaining`-th lowest candidate becomes the next digit of the threshold
__global__ void pick_kernel(EvictDev* d, int hi, int bits) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const unsigned long long rem = d->remaining;
  const int nbins = 1 << bits;
  unsigned long long cum = 0;
  int chosen = nbins - 1;
  for (int b = 0; b < nbins; ++b) {
    const unsigned long long c = d->hist[b];
    if (cum + c >= rem) {
      chosen = b;
      break;
    }
    cum += c;
  }
  d->prefix |= (unsigned long long)chosen << (hi - bits);
  d->remaining = rem > cum ? rem - cum : 0;
  for (int b = 0; b < kHistBins; ++b) d->hist[b] = 0;
}

// erase every live key whose score is below the threshold, plus `remaining` of the keys tied at it.  The slot goes
// straight back to EMPTY (score 0): chains that ran through its bucket are mended by repair_kernel.
__global__ void __launch_bounds__(kThreadsE)
evict_apply_kernel(TableView t, unsigned long long* __restrict__ sc, EvictDev* d) {
  __shared__ unsigned s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t cap = t.capacity();
  const unsigned long long tau = d->prefix, quota = d->remaining;
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsE + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsE) >> 5;
  unsigned cnt = 0;
  for (size_t base = warp0 * 32; base < cap; base += nwarps * 32) {
    const size_t s = base + lane;
    const bool lv = s < cap && live_key_e(t.keys[s]);
    const unsigned long long v = lv ? sc[s] : 0ull;
    bool go = lv && v < tau;
    const bool tie = lv && v == tau;
    const unsigned tb = __ballot_sync(kFull, tie);
    if (tb) {
      unsigned long long first = 0;
      if (lane == __ffs(tb) - 1) first = atomicAdd(&d->tie_ticket, (unsigned long long)__popc(tb));
      first = __shfl_sync(kFull, first, __ffs(tb) - 1);
      if (tie && first + __popc(tb & ((1u << lane) - 1u)) < quota) go = true;
    }
    if (go) {
      t.keys[s] = kEmptyKey;
      sc[s] = 0ull;
    }
    cnt += __popc(__ballot_sync(kFull, go));
  }
  if (lane == 0 && cnt) atomicAdd(&s_cnt, cnt);
  __syncthreads();
  if (threadIdx.x == 0 && s_cnt) {
    atomicAdd(&t.st->size, (unsigned long long)(-(long long)s_cnt));
    atomicAdd(&t.st->used, (unsigned long long)(-(long long)s_cnt));
    atomicAdd(&d->n_evicted, (unsigned long long)s_cnt);
  }
}

// Repair rounds.  A live key is UNREACHABLE when a bucket between its home bucket and the bucket it sits in has an
// EMPTY slot (a probe for it stops there).  One round = two kernels, each of which changes the set of EMPTY slots in
// ONE direction only, which is what makes the parallel pass safe without locks:
//   repair_move_kernel   every unreachable key is COPIED (ordinary find-or-claim: first free slot of its chain,
//                        which lies before its present slot; rows / optimizer slots / score travel along).  EMPTY
//                        slots only disappear here, so a key that is reachable stays reachable -- in particular a
//                        fresh copy is never picked up by another warp while its rows are still being written.
//                        A key judged unreachable from a stale view is simply FOUND by the claim probe and skipped.
//   repair_sweep_kernel  a displaced key that has an earlier match along its probe chain is a stale copy: its slot
//                        goes back to EMPTY (score 0).  EMPTY slots only appear here; a probe cut short by one of
//                        them reports "no earlier copy" and the stale copy survives until the next round.
// Freed slots may cut chains further on: the host repeats rounds until one moves nothing and erases nothing.
template <int VEC>
__global__ void __launch_bounds__(kThreadsE)
repair_move_kernel(TableView t, unsigned long long* __restrict__ sc, RowGeom g, RowGeom gslot, int n_planes,
                   EvictDev* d) {
  __shared__ unsigned s_moved, s_used;
  if (threadIdx.x == 0) {
    s_moved = 0;
    s_used = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t cap = t.capacity();
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsE + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsE) >> 5;
  for (size_t base = warp0 * 32; base < cap; base += nwarps * 32) {
    const size_t s = base + lane;
    const long long key = s < cap ? ld_key_cg(t.keys + s) : kEmptyKey;
    bool need = false;
    if (live_key_e(key)) {
      const unsigned long long bs = s / kBucket;
      unsigned long long b = bucket_of(key, t.nb);
      while (b != bs) {
        const long long* bp = t.keys + b * kBucket;
        bool has_empty = false;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const longlong2 kk = ld_keys_cg(bp + q * 2);
          has_empty |= (kk.x == kEmptyKey) | (kk.y == kEmptyKey);
        }
        if (has_empty) {
          need = true;
          break;
        }
        b = (b + 1 == t.nb) ? 0 : b + 1;
      }
    }
    if (!__any_sync(kFull, need)) continue;
    bool is_new, from_empty;
    const long long ns = warp_find_or_claim(t, key, need, need, lane, is_new, from_empty);
    const bool ok = need && ns >= 0 && is_new;
    warp_move_rows<VEC>(g, ok ? t.planes[0] + s * g.row_bytes : nullptr,
                        ok ? t.planes[0] + (size_t)ns * g.row_bytes : nullptr, lane);
    for (int p = 1; p <= n_planes; ++p)
      warp_move_rows<4>(gslot, ok ? t.planes[p] + s * gslot.row_bytes : nullptr,
                        ok ? t.planes[p] + (size_t)ns * gslot.row_bytes : nullptr, lane);
    if (ok) sc[ns] = sc[s];
    const unsigned bm = __ballot_sync(kFull, ok), bu = __ballot_sync(kFull, ok && from_empty);
    if (lane == 0 && bm) {
      atomicAdd(&s_moved, __popc(bm));
      atomicAdd(&s_used, __popc(bu));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_moved) {
    atomicAdd(&d->n_moved, (unsigned long long)s_moved);
    if (s_used) atomicAdd(&t.st->used, (unsigned long long)s_used);  // copies that consumed an EMPTY slot
  }
}

__global__ void __launch_bounds__(kThreadsE)
repair_sweep_kernel(TableView t, unsigned long long* __restrict__ sc, EvictDev* d) {
  __shared__ unsigned s_erased;
  if (threadIdx.x == 0) s_erased = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t cap = t.capacity();
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsE + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsE) >> 5;
  for (size_t base = warp0 * 32; base < cap; base += nwarps * 32) {
    const size_t s = base + lane;
    const long long key = s < cap ? ld_key_cg(t.keys + s) : kEmptyKey;
    const bool displaced = live_key_e(key) && bucket_of(key, t.nb) != s / kBucket;
    if (!__any_sync(kFull, displaced)) continue;
    const long long first = warp_find_slots<true>(t, key, displaced, lane);  // first match along the chain
    const bool stale = displaced && first >= 0 && (size_t)first != s;
    if (stale) {
      sc[s] = 0ull;
      st_key_cg(t.keys + s, kEmptyKey);
    }
    const unsigned be = __ballot_sync(kFull, stale);
    if (lane == 0 && be) atomicAdd(&s_erased, __popc(be));
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_erased) {
    atomicAdd(&d->n_erased, (unsigned long long)s_erased);
    atomicAdd(&t.st->used, (unsigned long long)(-(long long)s_erased));
  }
}

// growth: scores follow their keys into the new planes
__global__ void __launch_bounds__(kThreadsE)
carry_scores_kernel(TableView src, const unsigned long long* __restrict__ old_sc, TableView dst,
                    unsigned long long* __restrict__ new_sc) {
  const int lane = threadIdx.x & 31;
  const size_t cap = src.capacity();
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsE + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsE) >> 5;
  for (size_t base = warp0 * 32; base < cap; base += nwarps * 32) {
    const size_t s = base + lane;
    const long long key = s < cap ? src.keys[s] : kEmptyKey;
    const bool valid = live_key_e(key);
    if (!__any_sync(kFull, valid)) continue;
    const long long ns = warp_find_slots<true>(dst, key, valid, lane);
    if (valid && ns >= 0) new_sc[ns] = old_sc[s];
  }
}

// ================================================================================================
// Host side
// ================================================================================================
static ScoreRule rule_of(const det_table* t) {
  ScoreRule r;
  r.strategy = t->ev->strategy;
  r.epoch = t->ev->epoch;
  return r;
}

static inline unsigned long long lowmask(int bits) { return bits >= 64 ? ~0ull : ((1ull << bits) - 1ull); }

static det_status read_dev(det_table* t, cudaStream_t s) {
  CUDA_TRY(cudaMemcpyAsync(t->ev->h_dev, t->ev->dev, offsetof(EvictDev, hist), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return DET_OK;
}

static det_status read_state_e(det_table* t, cudaStream_t s, DevState* out) {
  CUDA_TRY(cudaMemcpyAsync(t->h_state, t->view.st, sizeof(DevState), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  *out = *t->h_state;
  return DET_OK;
}

template <typename F>
static det_status dispatch_vec_e(int vec, F&& f) {
  switch (vec) {
    case 16: return f(std::integral_constant<int, 16>());
    case 8: return f(std::integral_constant<int, 8>());
    case 4: return f(std::integral_constant<int, 4>());
    case 2: return f(std::integral_constant<int, 2>());
    default: return f(std::integral_constant<int, 1>());
  }
}

det_status evict_attach(det_table* t, int strategy) {
  if (strategy < DET_EVICT_LRU || strategy > DET_EVICT_CUSTOMIZED)
    return fail(DET_INVALID_ARGUMENT, "det_table_create: unknown eviction strategy in cfg.flags");
  if (t->external)
    return fail(DET_UNIMPLEMENTED, "det_table_create_in_region: eviction strategies are not available for tables in a caller-provided region");
  EvictState* ev = new EvictState();
  ev->strategy = strategy;
  const size_t n = t->view.capacity() + 2;
  cudaError_t e = cudaMalloc((void**)&ev->scores, n * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(ev->scores, 0, n * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMalloc((void**)&ev->dev, sizeof(EvictDev));
  if (e == cudaSuccess) e = cudaMemset(ev->dev, 0, sizeof(EvictDev));
  if (e == cudaSuccess) e = cudaMallocHost((void**)&ev->h_dev, sizeof(EvictDev));
  if (e != cudaSuccess) {
    if (ev->scores) cudaFree(ev->scores);
    if (ev->dev) cudaFree(ev->dev);
    if (ev->h_dev) cudaFreeHost(ev->h_dev);
    delete ev;
    cudaGetLastError();
    return fail(DET_OUT_OF_MEMORY, std::string("det_table_create: score plane: ") + cudaGetErrorString(e));
  }
  t->ev = ev;
  return DET_OK;
}

void evict_free(det_table* t) {
  if (!t->ev) return;
  if (t->ev->scores) cudaFree(t->ev->scores);
  if (t->ev->dev) cudaFree(t->ev->dev);
  if (t->ev->h_dev) cudaFreeHost(t->ev->h_dev);
  delete t->ev;
  t->ev = nullptr;
}

// rehash_to has enqueued the move of keys and rows from `ov` to `nv` on s: build the score plane of `nv`
det_status evict_on_rehash(det_table* t, const TableView& ov, const TableView& nv, cudaStream_t s) {
  EvictState* ev = t->ev;
  const size_t ocap = ov.capacity(), ncap = nv.capacity();
  unsigned long long* ns = nullptr;
  CUDA_TRY(cudaMalloc((void**)&ns, (ncap + 2) * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(ns, 0, (ncap + 2) * sizeof(unsigned long long), s));
  carry_scores_kernel<<<grid_for(ocap, kThreadsE, t->sm_count, 8), kThreadsE, 0, s>>>(ov, ev->scores, nv, ns);
  CUDA_TRY(cudaMemcpyAsync(ns + ncap, ev->scores + ocap, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(s));
  cudaFree(ev->scores);
  ev->scores = ns;
  return DET_OK;
}

void evict_on_clear(det_table* t, cudaStream_t s) {
  cudaMemsetAsync(t->ev->scores, 0, (t->view.capacity() + 2) * sizeof(unsigned long long), s);
  t->last_used_snap = 0;
}

det_status evict_before_remove(det_table* t, const long long* keys, size_t n, cudaStream_t s) {
  scores_of_keys_kernel<<<grid_for(n, kThreadsE, t->sm_count, 8), kThreadsE, 0, s>>>(t->view, keys, n, t->ev->scores,
                                                                                  nullptr, 1);
  CUDA_TRY(cudaGetLastError());
  return DET_OK;
}

det_status evict_touch(det_table* t, const long long* keys, const unsigned long long* scores, size_t n, cudaStream_t s) {
  touch_kernel<<<grid_for(n, kThreadsE, t->sm_count, 8), kThreadsE, 0, s>>>(t->view, keys, scores, n, t->ev->scores,
                                                                         rule_of(t));
  CUDA_TRY(cudaGetLastError());
  return DET_OK;
}

bool evict_at_max(const det_table* t) {
  if (t->cfg.max_capacity == 0) return false;  // scores are kept, the table grows without bound: det_evict only
  const uint64_t max_nb = (t->cfg.max_capacity + kBucket - 1) / kBucket;
  return t->view.nb >= max_nb;
}

// The eviction event: remove the k lowest-scored keys (all of them when fewer are resident).  Caller holds t->mu.
det_status evict_lowest(det_table* t, uint64_t k, cudaStream_t s) {
  EvictState* ev = t->ev;
  const TableView v = t->view;
  const size_t cap = v.capacity();
  const int grid = grid_for(cap, kThreadsE * 4, t->sm_count, 8);
  evict_reset_kernel<<<1, 256, 0, s>>>(ev->dev);
  minmax_kernel<<<grid, kThreadsE, 0, s>>>(v, ev->scores, ev->dev);
  CUDA_TRY(cudaGetLastError());
  det_status st = read_dev(t, s);
  if (st != DET_OK) return st;
  const unsigned long long n_live = ev->h_dev->n_live, smin = ev->h_dev->smin, smax = ev->h_dev->smax;
  if (n_live == 0 || k == 0) return DET_OK;
  if (k > n_live) k = n_live;
  int sig = 0;
  for (unsigned long long diff = smin ^ smax; diff; diff >>= 1) ++sig;   // bits in which the scores differ
  select_init_kernel<<<1, 1, 0, s>>>(ev->dev, smin & ~lowmask(sig), k);
  for (int hi = sig; hi > 0;) {
    const int bits = hi < kHistBits ? hi : kHistBits;
    hist_kernel<<<grid, kThreadsE, 0, s>>>(v, ev->scores, ev->dev, hi, bits);
    pick_kernel<<<1, 32, 0, s>>>(ev->dev, hi, bits);
    hi -= bits;
  }
  evict_apply_kernel<<<grid, kThreadsE, 0, s>>>(v, ev->scores, ev->dev);
  CUDA_TRY(cudaGetLastError());
  // repair rounds
  const int vec = pick_vec(t->row_bytes, nullptr, nullptr, nullptr);
  const RowGeom g = make_geom((unsigned)t->row_bytes, vec);
  const RowGeom gs = make_geom((unsigned)t->cfg.dim * 4u, 4);
  const int np = t->cfg.num_slot_planes;
  for (int round = 0; round < 256; ++round) {
    CUDA_TRY(cudaMemsetAsync(&ev->dev->n_moved, 0, 2 * sizeof(unsigned long long), s));  // n_moved, n_erased
    const int rgrid = grid_for(cap, kThreadsE, t->sm_count, 8);
    dispatch_vec_e(vec, [&](auto V) -> det_status {
      repair_move_kernel<decltype(V)::value><<<rgrid, kThreadsE, 0, s>>>(v, ev->scores, g, gs, np, ev->dev);
      return DET_OK;
    });
    repair_sweep_kernel<<<rgrid, kThreadsE, 0, s>>>(v, ev->scores, ev->dev);
    CUDA_TRY(cudaGetLastError());
    st = read_dev(t, s);
    if (st != DET_OK) return st;
    if (ev->h_dev->n_moved == 0 && ev->h_dev->n_erased == 0) break;
  }
  ev->n_events++;
  ev->n_evicted += ev->h_dev->n_evicted;
  return DET_OK;
}

// Room for a launch of n keys on a table that cannot grow any more (called by ensure_room under t->mu).
//  * steady state: no sync.  The load limit is a SOFT target: an event is due once the last landed snapshot of the
//    slot counter has reached it; the hard bound (95 % of the slots) is enforced on the host's upper bound.
//  * otherwise: exact counters, exact number of new keys of the batch (and, for det_insert_scored, which of them are
//    admitted), then one eviction event that frees what the batch needs plus a slab.
det_status evict_room(det_table* t, const long long* keys, size_t n, cudaStream_t s) {
  EvictState* ev = t->ev;
  ev->ctx_mask = nullptr;
  const uint64_t cap = t->view.capacity();
  const uint64_t limit = (uint64_t)((double)cap * t->max_lf);
  uint64_t hard = (uint64_t)((double)cap * 0.95);
  if (hard < limit) hard = limit;
  if (t->used_ub + n <= hard && t->last_used_snap < limit) {
    t->used_ub += n;
    return DET_OK;
  }
  DevState ds;
  det_status st = read_state_e(t, s, &ds);
  if (st != DET_OK) return st;
  t->snap_inflight = false;
  t->used_ub = ds.used;
  t->last_used_snap = ds.used;
  if (ds.used + n <= limit) {
    t->used_ub += n;
    return DET_OK;
  }
  uint64_t n_adm = n;
  unsigned char* mask = nullptr;
  if (keys != nullptr) {
    const bool admission = ev->ctx_admission;
    if (admission) {
      void* sc = nullptr;
      st = table_scratch(t, n, &sc);
      if (st != DET_OK) return st;
      mask = (unsigned char*)sc;
    }
    evict_reset_kernel<<<1, 256, 0, s>>>(ev->dev);
    if (admission)
      minmax_kernel<<<grid_for(cap, kThreadsE * 4, t->sm_count, 8), kThreadsE, 0, s>>>(t->view, ev->scores, ev->dev);
    classify_kernel<<<grid_for(n, kThreadsE, t->sm_count, 8), kThreadsE, 0, s>>>(t->view, keys, ev->ctx_scores, n, rule_of(t),
                                                                              admission ? 1 : 0, mask, ev->dev);
    CUDA_TRY(cudaGetLastError());
    st = read_dev(t, s);
    if (st != DET_OK) return st;
    n_adm = ev->h_dev->n_adm;
  }
  ev->ctx_mask = mask;
  if (ds.used + n_adm <= limit) {
    t->used_ub = ds.used + n_adm;
    return DET_OK;
  }
  const uint64_t special = ds.special[0] + ds.special[1];
  const uint64_t live = ds.size - special;
  const uint64_t need = ds.used + n_adm - limit;
  uint64_t slab = limit / 32;
  if (slab < 1) slab = 1;
  uint64_t k = need + slab;
  if (k > live) k = live;
  st = evict_lowest(t, k, s);
  if (st != DET_OK) return st;
  st = read_state_e(t, s, &ds);
  if (st != DET_OK) return st;
  t->last_used_snap = ds.used;
  if (ds.used + n_adm > hard)
    return fail(DET_TABLE_FULL, "detable: " + std::to_string(n_adm) + " new keys do not fit under max_capacity (" +
                                    std::to_string(cap) + " slots) even after eviction; send smaller batches");
  t->used_ub = ds.used + n_adm;
  return DET_OK;
}

// det_insert / det_insert_scored on a table with an eviction strategy
det_status evict_insert(det_table* t, const int64_t* keys, const void* values, const uint64_t* scores, size_t n,
                        cudaStream_t s) {
  std::lock_guard<std::mutex> _lk(t->mu);
  if (n == 0) return DET_OK;
  if (!keys || !values) return fail(DET_INVALID_ARGUMENT, "det_insert: null keys/values");
  EvictState* ev = t->ev;
  if (ev->strategy == DET_EVICT_CUSTOMIZED && scores == nullptr)
    return fail(DET_INVALID_ARGUMENT, "det_insert: the CUSTOMIZED eviction strategy needs scores (det_insert_scored)");
  det::DevGuard _dg(t->cfg.device);
  // a launch never brings more new keys than a quarter of what the table may hold (room is made per launch)
  const uint64_t max_cap = ((t->cfg.max_capacity + kBucket - 1) / kBucket) * kBucket;
  uint64_t chunk = t->cfg.max_capacity ? (uint64_t)((double)max_cap * t->max_lf) / 4 : (uint64_t)n;
  if (chunk < 1) chunk = 1;
  const int vec = pick_vec(t->row_bytes, values, nullptr, nullptr);
  const RowGeom g = make_geom((unsigned)t->row_bytes, vec);
  for (size_t off = 0; off < n; off += chunk) {
    const size_t m = (n - off < chunk) ? n - off : (size_t)chunk;
    const long long* k = (const long long*)keys + off;
    const unsigned long long* sc_in = scores ? (const unsigned long long*)scores + off : nullptr;
    const unsigned char* vals = (const unsigned char*)values + off * t->row_bytes;
    ev->ctx_scores = sc_in;
    ev->ctx_admission = true;
    ev->ctx_mask = nullptr;
    det_status st = ensure_room(t, k, m, s);
    ev->ctx_scores = nullptr;
    ev->ctx_admission = false;
    if (st != DET_OK) return st;
    const unsigned char* mask = ev->ctx_mask;
    ev->ctx_mask = nullptr;
    const TableView v = t->view;
    const ScoreRule rule = rule_of(t);
    const int np = t->cfg.num_slot_planes;
    st = dispatch_vec_e(vec, [&](auto V) -> det_status {
      constexpr int VV = decltype(V)::value;
      const int grid = grid_for(m, kThreadsE, t->sm_count, occupancy_of(insert_scored_kernel<VV>, kThreadsE));
      insert_scored_kernel<VV><<<grid, kThreadsE, 0, s>>>(v, k, vals, sc_in, mask, m, g, np, ev->scores, rule);
      CUDA_TRY(cudaGetLastError());
      return DET_OK;
    });
    if (st != DET_OK) return st;
    note_mutation(t, m, s);
  }
  return DET_OK;
}

}  // namespace det

using namespace det;

extern "C" {

det_status det_insert_scored(det_table* t, const int64_t* keys, const void* values, const uint64_t* scores, size_t n,
                             det_stream_t stream) {
  if (!t) return fail(DET_INVALID_ARGUMENT, "det_insert_scored: null table");
  if (!t->ev) {
    if (scores) return fail(DET_INVALID_ARGUMENT, "det_insert_scored: the table was created without an eviction strategy");
    return det_insert(t, keys, values, n, stream);
  }
  return evict_insert(t, keys, values, scores, n, (cudaStream_t)stream);
}

det_status det_find_scores(det_table* t, const int64_t* keys, size_t n, uint64_t* scores_out, det_stream_t stream) {
  if (!t) return fail(DET_INVALID_ARGUMENT, "det_find_scores: null table");
  if (!t->ev) return fail(DET_INVALID_ARGUMENT, "det_find_scores: the table was created without an eviction strategy");
  if (n == 0) return DET_OK;
  if (!keys || !scores_out) return fail(DET_INVALID_ARGUMENT, "det_find_scores: null argument");
  det::DevGuard _dg(t->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  scores_of_keys_kernel<<<grid_for(n, kThreadsE, t->sm_count, 8), kThreadsE, 0, s>>>(
      t->view, (const long long*)keys, n, t->ev->scores, (unsigned long long*)scores_out, 0);
  CUDA_TRY(cudaGetLastError());
  return DET_OK;
}

det_status det_set_global_epoch(det_table* t, uint64_t epoch) {
  if (!t) return fail(DET_INVALID_ARGUMENT, "det_set_global_epoch: null table");
  if (!t->ev) return fail(DET_INVALID_ARGUMENT, "det_set_global_epoch: the table was created without an eviction strategy");
  std::lock_guard<std::mutex> _lk(t->mu);
  t->ev->epoch = epoch;
  return DET_OK;
}

det_status det_evict(det_table* t, uint64_t n_evict, int64_t* n_evicted_out_host, det_stream_t stream) {
  if (!t) return fail(DET_INVALID_ARGUMENT, "det_evict: null table");
  if (!t->ev) return fail(DET_INVALID_ARGUMENT, "det_evict: the table was created without an eviction strategy");
  std::lock_guard<std::mutex> _lk(t->mu);
  det::DevGuard _dg(t->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  const uint64_t before = t->ev->n_evicted;
  det_status st = evict_lowest(t, n_evict, s);
  if (st != DET_OK) return st;
  if (n_evicted_out_host) *n_evicted_out_host = (int64_t)(t->ev->n_evicted - before);
  DevState ds;
  st = read_state_e(t, s, &ds);
  if (st != DET_OK) return st;
  t->snap_inflight = false;
  t->used_ub = ds.used;
  t->last_used_snap = ds.used;
  return DET_OK;
}

}  // extern "C"

namespace det {
void evict_stats(const det_table* t, uint32_t* events, uint64_t* evicted) {
  *events = t->ev ? t->ev->n_events : 0;
  *evicted = t->ev ? t->ev->n_evicted : 0;
}
}  // namespace det
