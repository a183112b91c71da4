// fused.cu -- kernels that replace CHAINS of reference ops on the hot path:
//   K6  det_lookup_sparse   unique -> find -> gather*weights -> segment_sum -> normalise
//                           (python/ops/dynamic_embedding_ops.py:219-291)
//   K7  det_apply_adagrad / det_apply_adam
//                           find(param)+find(slots) -> dense rule -> upsert(param)+upsert(slots)
//                           (python/ops/dynamic_embedding_optimizer.py:161-204)
//   K8  det_partition / det_gather_rows / det_scatter_rows
//                           default_partition_fn + dynamic_partition / dynamic_stitch
//                           (python/ops/dynamic_embedding_variable.py:131-197)
//       det_unique          tf.unique, first-occurrence order
// Compiled with --fmad=false: every fp32 multiply and add rounds separately, so results are
// bit-identical to the NumPy restatement in oracle/oracle.py.
#include "host.h"

namespace det {

constexpr int kThreadsF = 256;

// ------------------------------------------------------------------------------------------------
// generic exclusive scan of uint32 flags (n up to 2^32) : block sums -> single-block scan -> apply
// ------------------------------------------------------------------------------------------------
constexpr int kScanBlock = 1024;

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned* s_warp, unsigned& total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned x = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(kFull, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[w] = x;
  __syncthreads();
  if (w == 0) {
    const unsigned ws = s_warp[lane];
    unsigned xs = ws;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(kFull, xs, o);
      if (lane >= o) xs += y;
    }
    s_warp[lane] = xs - ws;
    if (lane == 31) s_warp[32] = xs;
  }
  __syncthreads();
  const unsigned excl = s_warp[w] + x - v;
  total = s_warp[32];
  __syncthreads();
  return excl;
}

__global__ void __launch_bounds__(kScanBlock)
scan_block_sums_kernel(const unsigned* __restrict__ flags, size_t n, unsigned* __restrict__ block_sums,
                       const long long* __restrict__ n_dev = nullptr) {
  if (n_dev) n = (size_t)*n_dev;   // item count produced on the device; the grid covers its bound
  __shared__ unsigned s_warp[33];
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  unsigned total;
  block_exclusive_scan(i < n ? flags[i] : 0u, s_warp, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: in-place exclusive scan of block sums (64-bit carry), total -> *total_out
__global__ void __launch_bounds__(kScanBlock)
scan_sums_kernel(unsigned* __restrict__ block_sums, size_t nblocks, long long* __restrict__ total_out) {
  __shared__ unsigned s_warp[33];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (size_t base = 0; base < nblocks; base += kScanBlock) {
    const size_t i = base + threadIdx.x;
    const unsigned v = i < nblocks ? block_sums[i] : 0u;
    unsigned total;
    const unsigned excl = block_exclusive_scan(v, s_warp, total);
    const unsigned long long carry = s_carry;
    if (i < nblocks) block_sums[i] = (unsigned)(carry + excl);
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = (long long)s_carry;
}

// ------------------------------------------------------------------------------------------------
// det_unique
// ------------------------------------------------------------------------------------------------
struct UniqueWs {
  long long* hkeys;     // [hcap]
  unsigned* hmin;       // [hcap] min position of the key, later its rank
  unsigned* myslot;     // [n]
  unsigned* flags;      // [n]
  unsigned* block_sums; // [nblocks]
  unsigned* special;    // [2]: min position / rank of the key equal to the scratch sentinel
  size_t hcap;          // power of two
};

static size_t unique_hcap(size_t n) {
  size_t c = 64;
  while (c < 2 * n) c <<= 1;
  return c;
}

__global__ void unique_init_kernel(UniqueWs w) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < w.hcap; i += (size_t)gridDim.x * blockDim.x) {
    w.hkeys[i] = kEmptyKey;
    w.hmin[i] = 0xffffffffu;
  }
  if (blockIdx.x == 0 && threadIdx.x < 2) w.special[threadIdx.x] = 0xffffffffu;
}

__global__ void unique_insert_kernel(UniqueWs w, const long long* __restrict__ ids, size_t n,
                                     const long long* __restrict__ n_dev) {
  if (n_dev) n = (size_t)*n_dev;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long key = ids[i];
  if (key == kEmptyKey) {
    atomicMin(&w.special[0], (unsigned)i);
    w.myslot[i] = 0xffffffffu;
    return;
  }
  size_t s = fmix64((unsigned long long)key) & (w.hcap - 1);
  while (true) {
    const long long old = (long long)atomicCAS((unsigned long long*)(w.hkeys + s), (unsigned long long)kEmptyKey,
                                               (unsigned long long)key);
    if (old == kEmptyKey || old == key) {
      atomicMin(&w.hmin[s], (unsigned)i);
      w.myslot[i] = (unsigned)s;
      return;
    }
    s = (s + 1) & (w.hcap - 1);
  }
}

__global__ void unique_flag_kernel(UniqueWs w, size_t n, const long long* __restrict__ n_dev) {
  if (n_dev) n = (size_t)*n_dev;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned s = w.myslot[i];
  const unsigned first = (s == 0xffffffffu) ? w.special[0] : w.hmin[s];
  w.flags[i] = (first == (unsigned)i) ? 1u : 0u;
}

// unique_flag_kernel + scan_block_sums_kernel in one pass: flag of every position and the flag count of its block
__global__ void __launch_bounds__(kScanBlock)
unique_flag_sums_kernel(UniqueWs w, size_t n, const long long* __restrict__ n_dev) {
  if (n_dev) n = (size_t)*n_dev;
  __shared__ unsigned s_warp[33];
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  unsigned f = 0u;
  if (i < n) {
    const unsigned s = w.myslot[i];
    const unsigned first = (s == 0xffffffffu) ? w.special[0] : w.hmin[s];
    f = (first == (unsigned)i) ? 1u : 0u;
    w.flags[i] = f;
  }
  unsigned total;
  block_exclusive_scan(f, s_warp, total);
  if (threadIdx.x == 0) w.block_sums[blockIdx.x] = total;
}

// first occurrences: rank = exclusive scan; write unique_out[rank], remember rank in the hash slot
__global__ void __launch_bounds__(kScanBlock)
unique_rank_kernel(UniqueWs w, const long long* __restrict__ ids, size_t n, long long* __restrict__ unique_out,
                   const long long* __restrict__ n_dev) {
  if (n_dev) n = (size_t)*n_dev;
  __shared__ unsigned s_warp[33];
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  const unsigned f = i < n ? w.flags[i] : 0u;
  unsigned total;
  const unsigned excl = block_exclusive_scan(f, s_warp, total);
  if (i < n && f) {
    const unsigned rank = w.block_sums[blockIdx.x] + excl;
    unique_out[rank] = ids[i];
    const unsigned s = w.myslot[i];
    if (s == 0xffffffffu) w.special[1] = rank; else w.hmin[s] = rank;
  }
}

__global__ void unique_idx_kernel(UniqueWs w, size_t n, int* __restrict__ idx_out, const long long* __restrict__ n_dev) {
  if (n_dev) n = (size_t)*n_dev;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned s = w.myslot[i];
  idx_out[i] = (int)((s == 0xffffffffu) ? w.special[1] : w.hmin[s]);
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t unique_ws_layout(size_t n, unsigned char* base, UniqueWs* w) {
  const size_t hcap = unique_hcap(n);
  const size_t nblocks = (n + kScanBlock - 1) / kScanBlock;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  long long* hkeys = (long long*)take(hcap * 8);
  unsigned* hmin = (unsigned*)take(hcap * 4);
  unsigned* myslot = (unsigned*)take(n * 4);
  unsigned* flags = (unsigned*)take(n * 4);
  unsigned* bs = (unsigned*)take((nblocks + 1) * 4);
  unsigned* sp = (unsigned*)take(16);
  if (w) {
    w->hkeys = hkeys; w->hmin = hmin; w->myslot = myslot; w->flags = flags;
    w->block_sums = bs; w->special = sp; w->hcap = hcap;
  }
  return off;
}

// ------------------------------------------------------------------------------------------------
// K6: fused embedding_lookup_sparse
// ------------------------------------------------------------------------------------------------
// seg_start[b] = first position i with segment_ids[i] >= b ; seg_start[batch] = nnz
// run_if_set (nullable): the flag lookup_identity_kernel leaves behind -- zero means "exactly one id per output row was
// verified and served", the Criteo shape; the general kernels then have nothing to do
__global__ void segment_offsets_kernel(const int* __restrict__ seg, size_t nnz, size_t batch,
                                       long long* __restrict__ seg_start, DevState* st, const unsigned* run_if_set,
                                       unsigned epoch) {
  if (run_if_set && *run_if_set != epoch) return;   // one id per row, verified by lookup_identity_kernel: nothing to do
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nnz) return;
  const long long cur = i < nnz ? (long long)seg[i] : (long long)batch;
  const long long prev = i > 0 ? (long long)seg[i - 1] : -1;
  if (cur < prev || cur < 0 || cur > (long long)batch || (i < nnz && cur >= (long long)batch)) {
    atomicOr(&st->error, kErrBadSegment);
    return;
  }
  for (long long b = prev + 1; b <= cur; ++b) seg_start[b] = (long long)i;
}

// K6 phase A: slot of every id (-1 = absent), 32 ids per warp-step with the same warp-cooperative probe as
// det_find; only 8 B per id leave the kernel (the [nnz, dim] gather of the reference is never materialised)
__global__ void __launch_bounds__(kThreadsF)
resolve_slots_kernel(TableView t, const long long* __restrict__ ids, size_t nnz, long long* __restrict__ slots,
                     int use_tma, const unsigned* __restrict__ run_if_set, unsigned epoch) {
  if (run_if_set && *run_if_set != epoch) return;   // the one-id-per-row kernel has served this call
  __shared__ __align__(128) long long s_keys[kStages][kTileKeys];
  __shared__ __align__(8) unsigned long long s_bar[kStages];
  const int lane = threadIdx.x & 31;
  KeyTiles kt;
  kt.init(s_keys, s_bar, ids, nnz, use_tma != 0);
  for (; kt.valid(); kt.next()) {
    size_t i;
    bool valid;
    const long long key = kt.key(i, valid);
    const long long slot = warp_find_slots<false>(t, key, valid, lane);
    if (valid) slots[i] = slot;
  }
}

// K6, one id per output row and no weights (every combiner then returns the row itself: 0 + row * 1, / 1): the call IS a
// Find into the dense output, so it runs as one -- 32 rows per warp-step, 4 row loads in flight per lane, no slot
// scratch and no per-row segment bookkeeping.  Launched on spec IN FRONT of the general kernels (nnz == batch, no weights):
// it checks segment_ids[i] == i for the ids it serves and raises `not_identity` otherwise; the general kernels
// (segment_offsets, resolve_slots, segment_sum) run only when the flag is set and then rewrite every output row.
template <int VEC>
__global__ void __launch_bounds__(kThreadsF)
lookup_identity_kernel(TableView t, const long long* __restrict__ ids, const int* __restrict__ seg, size_t n,
                       const unsigned char* __restrict__ default_row, unsigned char* __restrict__ out, RowGeom g, int use_tma,
                       unsigned* __restrict__ not_identity, unsigned epoch) {
  __shared__ __align__(128) long long s_keys[kStages][kTileKeys];
  __shared__ __align__(8) unsigned long long s_bar[kStages];
  const int lane = threadIdx.x & 31;
  KeyTiles kt;
  kt.init(s_keys, s_bar, ids, n, use_tma != 0);
  for (; kt.valid(); kt.next()) {
    size_t i;
    bool valid;
    const long long key = kt.key(i, valid);
    const bool off_row = valid && __ldg(seg + i) != (int)i;        // this id does not belong to output row i
    if (__any_sync(kFull, off_row) && lane == 0) *not_identity = epoch;   // "raised in THIS call": the flag is never reset
    const long long slot = warp_find_slots<false>(t, key, valid, lane);
    const unsigned char* src = nullptr;
    unsigned char* dst = nullptr;
    if (valid) {
      src = slot >= 0 ? t.planes[0] + (size_t)slot * g.row_bytes : DET_SRC_DEFAULT;
      dst = out + i * g.row_bytes;
    }
    warp_move_rows<VEC>(g, src, dst, lane, default_row);
  }
}

template <int VF> struct FVec;
template <> struct FVec<4> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void fill(float x) { v = make_float4(x, x, x, x); }
  template <typename F> __device__ __forceinline__ void apply(F f) { f(v.x); f(v.y); f(v.z); f(v.w); }
  template <typename F> __device__ __forceinline__ void zip(const FVec& o, F f) { f(v.x, o.v.x); f(v.y, o.v.y); f(v.z, o.v.z); f(v.w, o.v.w); }
};
template <> struct FVec<1> {
  float v;
  __device__ __forceinline__ void load(const float* p) { v = *p; }
  __device__ __forceinline__ void store(float* p) const { *p = v; }
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void fill(float x) { v = x; }
  template <typename F> __device__ __forceinline__ void apply(F f) { f(v); }
  template <typename F> __device__ __forceinline__ void zip(const FVec& o, F f) { f(v, o.v); }
};

constexpr int kMaxVecPerLane = 8;
// segments a lane-group works on concurrently (independent load chains) and the resident CTAs per SM the register
// budget is capped for; measurement builds override them (-DDET_SEG_U=8 -DDET_SEG_MINB=2, scripts/segsum_sweep.sh)
#ifndef DET_SEG_U
#define DET_SEG_U 4
#endif
#ifndef DET_SEG_MINB
#define DET_SEG_MINB 3
#endif
constexpr int kSegPerGroup = DET_SEG_U;

// K6 phase B, common case (one vector per lane covers the row): each lane-group owns kSegPerGroup
// consecutive output rows and walks their ids in lock step, so up to 4 independent row loads are in flight
// per lane; within a segment the ids are accumulated strictly in order (mul, then add: the summation order of
// the reference test oracle).
// max_norm of embedding_lookup(_sparse) (tf.clip_by_norm over each looked-up row, python/ops/embedding_weights.py:497-521)
// folded into the gather: an EMPTY trailing kernel parameter when unused, so the unclipped kernel is unchanged.
template <bool CLIP> struct ClipArg {};
template <> struct ClipArg<true> { float max_norm; };
template <bool CLIP> __device__ __forceinline__ float clip_norm_of(const ClipArg<CLIP>&) { return 0.f; }
template <> __device__ __forceinline__ float clip_norm_of<true>(const ClipArg<true>& c) { return c.max_norm; }

template <int VF, bool CLIP = false>
__global__ void __launch_bounds__(kThreadsF, DET_SEG_MINB)
segment_sum_kernel(TableView t, const long long* __restrict__ slots, const long long* __restrict__ seg_start,
                   const float* __restrict__ weights, size_t batch, int combiner,
                   const float* __restrict__ default_row, float* __restrict__ out, unsigned vpr, unsigned lpr,
                   unsigned lpr_shift, ClipArg<CLIP> clip, const unsigned* __restrict__ run_if_set, unsigned epoch) {
  if (run_if_set && *run_if_set != epoch) return;   // the one-id-per-row kernel has served this call
  constexpr int U = kSegPerGroup;
  const int lane = threadIdx.x & 31;
  const unsigned gl = (unsigned)lane & (lpr - 1u);
  const unsigned gpw = 32u >> lpr_shift;
  const unsigned gidx = (unsigned)lane >> lpr_shift;
  const unsigned dim = t.dim;
  const bool lane_on = gl < vpr;
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsF + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsF) >> 5;
  const float* table = (const float*)t.planes[0];
  FVec<VF> defv;
  defv.zero();
  if (lane_on) defv.load(default_row + (size_t)gl * VF);
  for (size_t sb = warp0 * gpw * U; sb < batch; sb += nwarps * gpw * U) {
    const size_t b0 = sb + (size_t)gidx * U;
    long long st[U], en[U];
    FVec<VF> acc[U];
    float wsum[U], wsq[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = b0 + u < batch;
      st[u] = ok ? seg_start[b0 + u] : 0;
      en[u] = ok ? seg_start[b0 + u + 1] : 0;
      acc[u].zero();
      wsum[u] = 0.f;
      wsq[u] = 0.f;
    }
    for (long long k = 0;; ++k) {
      bool any = false;
#pragma unroll
      for (int u = 0; u < U; ++u) any |= st[u] + k < en[u];
      if (!__any_sync(kFull, any)) break;
      long long sl[U];
      float w[U];
      FVec<VF> x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool act = st[u] + k < en[u];
        sl[u] = act ? __ldg(slots + st[u] + k) : -2;
        w[u] = act ? (weights ? __ldg(weights + st[u] + k) : 1.f) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        x[u] = defv;
        if (sl[u] >= 0 && lane_on) x[u].load(table + (size_t)sl[u] * dim + (size_t)gl * VF);
      }
      if (CLIP) {
        // row <- row * max_norm / max(||row||, max_norm): the squared norm is summed over the lanes of the group
        // (lanes beyond the row hold zeros); every lane of the warp takes part in the shuffles (uniform loop)
        const float mx = clip_norm_of(clip);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float ss = 0.f;
          x[u].apply([&ss](float& a) { ss = ss + a * a; });
          for (unsigned o = lpr >> 1; o > 0; o >>= 1) ss = ss + __shfl_xor_sync(kFull, ss, (int)o);
          const float scale = mx / fmaxf(sqrtf(ss), mx);
          x[u].apply([scale](float& a) { a = a * scale; });
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (sl[u] != -2) {
          const float wu = w[u];
          acc[u].zip(x[u], [wu](float& a, float xv) { a = a + xv * wu; });
          wsum[u] = wsum[u] + wu;
          wsq[u] = wsq[u] + wu * wu;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (b0 + u < batch && lane_on) {
        if (en[u] > st[u] && combiner != DET_COMBINER_SUM) {
          const float div = combiner == DET_COMBINER_MEAN ? wsum[u] : sqrtf(wsq[u]);
          acc[u].apply([div](float& a) { a = a / div; });
        }
        acc[u].store(out + (b0 + u) * dim + (size_t)gl * VF);
      }
    }
  }
}

// K6 phase B, staged variant (DET_SEGSUM_STAGED=1; a round-2 candidate, default off): the register-held kernel above
// keeps kSegPerGroup 16 B loads in flight per lane and cannot go further (8 segments spill or drop to one CTA per SM,
// scripts/segsum_sweep.sh); ncu showed it at 4.2 TB/s with nothing saturated.  Here every warp owns a CONTIGUOUS range
// of output rows, so the ids it needs are one contiguous stretch of `slots`: the warp streams that stretch through a
// 3-stage ring of shared memory in windows of kWinSteps row-steps with cp.async (LDGSTS: up to two windows = 4 KB per
// warp in flight, no registers held), and sums window k while k+1 and k+2 are loading.  A segment always belongs to the
// same lane-group (segment index mod groups per warp), which adds its rows strictly in id order and carries the partial
// sum across windows -- the arithmetic, its order and the results are those of segment_sum_kernel.
// Requires VF == 4 rows of one vector per lane (dim % 4 == 0, dim <= 128).
constexpr int kWinSteps = 4;   // row-steps per window: 4 x 32 lanes x 16 B = 2 KB
constexpr int kWinStages = 3;
constexpr int kWinFloat4 = kWinSteps * 32;                       // float4 per stage
constexpr int kWinStageBytes = kWinFloat4 * 16 + 32 * 4;         // rows + one weight per staged row
constexpr int kWinWarpBytes = kWinStages * kWinStageBytes;

__global__ void __launch_bounds__(kThreadsF)
segment_sum_staged_kernel(TableView t, const long long* __restrict__ slots, const long long* __restrict__ seg_start,
                          const float* __restrict__ weights, size_t batch, int combiner,
                          const float* __restrict__ default_row, float* __restrict__ out, unsigned vpr, unsigned lpr,
                          unsigned lpr_shift) {
  DET_DYN_SHARED(dyn_smem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned gl = (unsigned)lane & (lpr - 1u);
  const unsigned gpw = 32u >> lpr_shift;          // lane-groups per warp = rows per row-step
  const unsigned gidx = (unsigned)lane >> lpr_shift;
  const unsigned gsh = 5u - lpr_shift;            // log2(gpw)
  const unsigned wsteps = gpw * (unsigned)kWinSteps > 32u ? 32u / gpw : (unsigned)kWinSteps;
  const unsigned W = wsteps * gpw;                // rows per window (<= 32: lane r holds the slot of row r)
  const unsigned dim = t.dim;
  const bool lane_on = gl < vpr;
  const float* table = (const float*)t.planes[0];
  unsigned char* wbase = dyn_smem + (size_t)warp * kWinWarpBytes;
  auto s_row = [&](int stage) -> float4* { return reinterpret_cast<float4*>(wbase + (size_t)stage * kWinStageBytes); };
  auto s_wt = [&](int stage) -> float* { return reinterpret_cast<float*>(wbase + (size_t)stage * kWinStageBytes + kWinFloat4 * 16); };
  // this warp's contiguous range of segments [S0, S1)
  const size_t nwarps = ((size_t)gridDim.x * kThreadsF) >> 5;
  const size_t spw = (batch + nwarps - 1) / nwarps;
  const size_t S0 = ((((size_t)blockIdx.x * kThreadsF) >> 5) + (size_t)warp) * spw;
  if (S0 >= batch) return;
  const size_t S1 = S0 + spw < batch ? S0 + spw : batch;
  const unsigned nseg = (unsigned)(S1 - S0);
  const long long p_begin = seg_start[S0], p_end = seg_start[S1];
  const unsigned npos = (unsigned)(p_end - p_begin);
  const unsigned nwin = npos ? (npos + W - 1) / W : 1u;   // an empty range still flushes its (empty) segments
  // segment boundaries relative to p_begin, 64 at a time: lane i of b_lo holds boundary c*32 + i, b_hi the next 32
  auto load_bounds = [&](unsigned first) -> unsigned {
    const unsigned i = first + (unsigned)lane;
    return (unsigned)(seg_start[S0 + (i < nseg ? i : nseg)] - p_begin);
  };
  unsigned chunk = 0;
  unsigned b_lo = load_bounds(0), b_hi = load_bounds(32);
  auto bound = [&](unsigned i) -> unsigned {    // boundary i, chunk*32 <= i < chunk*32 + 64; warp-uniform call
    const unsigned o = i - chunk * 32u;
    const unsigned lo = __shfl_sync(kFull, b_lo, (int)(o & 31u)), hi = __shfl_sync(kFull, b_hi, (int)(o & 31u));
    return o < 32u ? lo : hi;
  };
  // producer: the slot of row `lane` of a window, then the cp.async copies of the window
  auto load_slot = [&](unsigned k) -> long long {
    const unsigned r = k * W + (unsigned)lane;
    return (k < nwin && (unsigned)lane < W && r < npos) ? __ldg(slots + p_begin + r) : -1;
  };
  auto issue = [&](unsigned k, long long sl_of_lane) {
    const int stage = (int)(k % (unsigned)kWinStages);
    const unsigned w0 = k * W;
#pragma unroll
    for (int stp = 0; stp < kWinSteps; ++stp) {
      const unsigned r = (unsigned)stp * gpw + gidx;
      const long long sl = shfl_ll(sl_of_lane, (int)r);
      if ((unsigned)stp < wsteps && k < nwin && w0 + r < npos && lane_on)
        cp_async16(s_row(stage) + stp * 32 + lane,
                   sl >= 0 ? table + (size_t)sl * dim + (size_t)gl * 4 : default_row + (size_t)gl * 4);
    }
    if (weights && k < nwin && (unsigned)lane < W && w0 + (unsigned)lane < npos)
      cp_async4(s_wt(stage) + lane, weights + p_begin + w0 + lane);
    cp_async_commit();
  };
  long long sl_next = load_slot(0);
  issue(0, sl_next);
  sl_next = load_slot(1);
  issue(1, sl_next);
  sl_next = load_slot(2);
  // consumer state: cs = first segment (relative) that is not finished; my group's running sums
  unsigned cs = 0;
  FVec<4> acc;
  acc.zero();
  float wsum = 0.f, wsq = 0.f;
  for (unsigned k = 0; k < nwin; ++k) {
    const long long sl_issue = sl_next;
    sl_next = load_slot(k + 3);                   // in flight while this window is summed
    issue(k + 2, sl_issue);
    cp_async_wait<2>();
    __syncwarp();
    const int stage = (int)(k % (unsigned)kWinStages);
    const unsigned w0 = k * W;
    const unsigned w1 = w0 + W < npos ? w0 + W : npos;
    const float4* rows = s_row(stage);
    const float* wts = s_wt(stage);
    unsigned done = cs;                           // segments [cs, done) are finished after this window
    if ((cs & ~(gpw - 1u)) < chunk * 32u) {       // the unfinished segment lies before the boundaries held: step back
      chunk = (cs & ~(gpw - 1u)) >> 5;
      b_lo = load_bounds(chunk * 32u);
      b_hi = load_bounds(chunk * 32u + 32u);
    }
    for (unsigned base = cs & ~(gpw - 1u); base < nseg; base += gpw) {
      while (base + gpw > chunk * 32u + 63u) {    // keep boundaries base .. base + gpw inside the 64 held
        ++chunk;
        b_lo = b_hi;
        b_hi = load_bounds(chunk * 32u + 32u);
      }
      const unsigned first_st = bound(base < cs ? cs : base);
      if (first_st > w1) break;                   // every later segment starts after this window (uniform)
      const unsigned seg = base + gidx;
      const unsigned st = bound(seg < nseg ? seg : nseg), en = bound(seg + 1u < nseg ? seg + 1u : nseg);
      if (seg >= cs && seg < nseg && st <= w1) {
        const unsigned a = st > w0 ? st : w0, b = en < w1 ? en : w1;
        for (unsigned p = a; p < b; ++p) {
          const unsigned r = p - w0;
          const float wu = weights ? wts[r] : 1.f;
          if (lane_on) {
            FVec<4> x;
            x.v = rows[(r >> gsh) * 32u + ((r & (gpw - 1u)) << lpr_shift) + gl];
            acc.zip(x, [wu](float& av, float xv) { av = av + xv * wu; });
          }
          wsum = wsum + wu;
          wsq = wsq + wu * wu;
        }
        if (en <= w1) {                           // the segment ends inside this window: normalise, store, reset
          if (lane_on) {
            if (en > st && combiner != DET_COMBINER_SUM) {
              const float div = combiner == DET_COMBINER_MEAN ? wsum : sqrtf(wsq);
              acc.apply([div](float& av) { av = av / div; });
            }
            acc.store(out + (S0 + seg) * dim + (size_t)gl * 4);
          }
          acc.zero();
          wsum = 0.f;
          wsq = 0.f;
          done = seg + 1u;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned other = __shfl_xor_sync(kFull, done, o);
      done = other > done ? other : done;
    }
    cs = done;
    __syncwarp();   // every lane is done with this stage before the next issue overwrites it
  }
  cp_async_wait<0>();
}

// K6 phase B, wide rows (several vectors per lane): one lane-group per output row
template <int VF>
__global__ void __launch_bounds__(kThreadsF)
segment_sum_wide_kernel(TableView t, const long long* __restrict__ slots, const long long* __restrict__ seg_start,
                        const float* __restrict__ weights, size_t batch, int combiner,
                        const float* __restrict__ default_row, float* __restrict__ out, unsigned vpr, unsigned lpr,
                        unsigned lpr_shift) {
  const int lane = threadIdx.x & 31;
  const int gl = lane & (int)(lpr - 1);
  const unsigned gpw = 32u >> lpr_shift;
  const unsigned gidx = (unsigned)lane >> lpr_shift;
  const unsigned dim = t.dim;
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsF + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsF) >> 5;
  const float* table = (const float*)t.planes[0];
  for (size_t sb = warp0 * gpw; sb < batch; sb += nwarps * gpw) {
    const size_t b = sb + gidx;
    const bool act_seg = b < batch;
    const long long start = act_seg ? seg_start[b] : 0;
    const long long end = act_seg ? seg_start[b + 1] : 0;
    FVec<VF> acc[kMaxVecPerLane];
#pragma unroll
    for (int v = 0; v < kMaxVecPerLane; ++v) acc[v].zero();
    float wsum = 0.f, wsq = 0.f;
    for (long long i = start; i < end; ++i) {
      const long long slot = __ldg(slots + i);
      const float w = weights ? __ldg(weights + i) : 1.f;
      const float* src = slot >= 0 ? table + (size_t)slot * dim : default_row;
#pragma unroll
      for (int v = 0; v < kMaxVecPerLane; ++v) {
        const unsigned c = (unsigned)gl + (unsigned)v * lpr;
        if (c < vpr) {
          FVec<VF> x;
          x.load(src + (size_t)c * VF);
          acc[v].zip(x, [w](float& a, float xv) { a = a + xv * w; });
        }
      }
      wsum = wsum + w;
      wsq = wsq + w * w;
    }
    if (act_seg) {
      float div = 1.f;
      bool do_div = false;
      if (end > start) {
        if (combiner == DET_COMBINER_MEAN) { div = wsum; do_div = true; }
        else if (combiner == DET_COMBINER_SQRTN) { div = sqrtf(wsq); do_div = true; }
      }
      float* dst = out + b * dim;
#pragma unroll
      for (int v = 0; v < kMaxVecPerLane; ++v) {
        const unsigned c = (unsigned)gl + (unsigned)v * lpr;
        if (c < vpr) {
          if (do_div) acc[v].apply([div](float& a) { a = a / div; });
          acc[v].store(dst + (size_t)c * VF);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K7: fused find-or-insert + optimizer update
// ------------------------------------------------------------------------------------------------
struct OptHyper {
  float lr;       // adagrad lr | adam alpha
  float eps;
  float beta1, beta2;
  float init_slot;  // adagrad initial accumulator
};

template <int VF, int OPT, int RU>
__global__ void __launch_bounds__(kThreadsF)
apply_kernel(TableView t, const long long* __restrict__ keys, const float* __restrict__ grads, size_t n,
             OptHyper h, const float* __restrict__ init_param, int full_init, unsigned vpr, unsigned lpr,
             unsigned lpr_shift, int use_tma) {
  __shared__ __align__(128) long long s_keys[kStages][kTileKeys];
  __shared__ __align__(8) unsigned long long s_bar[kStages];
  __shared__ unsigned s_new, s_used;
  if (threadIdx.x == 0) {
    s_new = 0;
    s_used = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned dim = t.dim;
  const unsigned rows_per_step = 32u >> lpr_shift;
  const unsigned sub = (unsigned)lane >> lpr_shift;
  const unsigned c0 = (unsigned)lane & (lpr - 1u);
  float* P = (float*)t.planes[0];
  float* S1 = (float*)t.planes[1];
  float* S2 = (float*)t.planes[2];
  const float omb1 = 1.f - h.beta1, omb2 = 1.f - h.beta2;
  KeyTiles kt;
  kt.init(s_keys, s_bar, keys, n, use_tma != 0);
  for (; kt.valid(); kt.next()) {
    size_t i;
    bool valid;
    const long long key = kt.key(i, valid);
    const size_t base = i - (size_t)lane;
    bool is_new, from_empty;
    const long long slot = warp_find_or_claim(t, key, valid, valid, lane, is_new, from_empty);
    const unsigned bn = __ballot_sync(kFull, is_new), bu = __ballot_sync(kFull, from_empty);
    if (lane == 0 && bn) {
      atomicAdd(&s_new, __popc(bn));
      atomicAdd(&s_used, __popc(bu));
    }
    // two row-steps per iteration: all loads of both rows are issued before the first update/store
    for (unsigned j0 = 0; j0 < 32u; j0 += rows_per_step * RU) {
      bool ok[RU], nwv[RU], fresh[RU];
      size_t ro[RU], gi[RU];
      unsigned mark[RU];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const unsigned j = j0 + (unsigned)u * rows_per_step + sub;
        const unsigned jj = j & 31u;
        const long long s = shfl_ll(slot, (int)jj);
        nwv[u] = (bn >> jj) & 1u;
        gi[u] = base + jj;
        ok[u] = (j < 32u) && (gi[u] < n) && s >= 0;
        ro[u] = ok[u] ? (size_t)s * dim : 0;
        // slot state absent: key created in this launch, or created by insert/accum and never stepped.
        // Every lane reads the marker word BEFORE any lane of the row's group overwrites it.
        mark[u] = ok[u] ? __float_as_uint(S1[ro[u]]) : 0u;
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < RU; ++u) fresh[u] = nwv[u] || (mark[u] == kSlotUninit);
      for (unsigned c = c0; c < vpr; c += lpr) {
        const size_t o = (size_t)c * VF;
        FVec<VF> g[RU], p[RU], a[RU], b[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          if (!ok[u]) continue;
          g[u].load(grads + gi[u] * dim + o);
          if (nwv[u]) p[u].load((full_init ? init_param + gi[u] * dim : init_param) + o);
          else p[u].load(P + ro[u] + o);
          if (OPT == 0) {
            if (fresh[u]) a[u].fill(h.init_slot); else a[u].load(S1 + ro[u] + o);
          } else {
            if (fresh[u]) { a[u].zero(); b[u].zero(); } else { a[u].load(S1 + ro[u] + o); b[u].load(S2 + ro[u] + o); }
          }
        }
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          if (!ok[u]) continue;
          if (OPT == 0) {
            // accum += g*g ; var -= lr*g / (sqrt(accum) + eps)
            a[u].zip(g[u], [](float& av, float gv) { av = av + gv * gv; });
            FVec<VF> upd = g[u];
            upd.zip(a[u], [&h](float& x, float av) { x = (h.lr * x) / (sqrtf(av) + h.eps); });
            p[u].zip(upd, [](float& pv, float x) { pv = pv - x; });
            a[u].store(S1 + ro[u] + o);
            p[u].store(P + ro[u] + o);
          } else {
            // m += (g-m)(1-b1) ; v += (g*g-v)(1-b2) ; var -= (m*alpha)/(sqrt(v)+eps)
            a[u].zip(g[u], [omb1](float& mv, float gv) { mv = mv + (gv - mv) * omb1; });
            b[u].zip(g[u], [omb2](float& vv, float gv) { vv = vv + (gv * gv - vv) * omb2; });
            FVec<VF> upd = a[u];
            upd.zip(b[u], [&h](float& x, float vv) { x = (x * h.lr) / (sqrtf(vv) + h.eps); });
            p[u].zip(upd, [](float& pv, float x) { pv = pv - x; });
            a[u].store(S1 + ro[u] + o);
            b[u].store(S2 + ro[u] + o);
            p[u].store(P + ro[u] + o);
          }
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_new) {
    atomicAdd(&t.st->size, (unsigned long long)s_new);
    atomicAdd(&t.st->used, (unsigned long long)s_used);
  }
}

// K7 variant: rows staged through shared memory with cp.async (LDGSTS), two batches in flight per warp.
// The plain kernel is latency-bound (ncu: DRAM 45 %, issue 35 %, nothing saturated): a lane holds the three
// (four) 16 B loads of ONE row-step in registers and stalls on them.  Here a lane's loads of the NEXT batch of
// row-steps are in flight (no registers held) while it updates the current batch.  Requires VF == 4 rows with one
// vector per lane (vpr <= lpr, dim <= 128).  Same arithmetic, same results.
constexpr int kStageSteps = 2;  // row-steps per batch

// Tables with an eviction strategy: the step also refreshes the score of every key it touches (HKV scores keys on
// find_or_insert / assign).  The slot is known right after the probe, so the score write is folded in here instead of
// a second probe pass (touch_kernel); an EMPTY trailing parameter when unused keeps the unscored kernel unchanged.
template <bool SCORED> struct ScoreArg {};
template <> struct ScoreArg<true> {
  unsigned long long* sc;
  ScoreRule rule;
};
template <bool SCORED>
__device__ __forceinline__ void score_touch(const ScoreArg<SCORED>&, long long, bool, bool) {}
template <>
__device__ __forceinline__ void score_touch<true>(const ScoreArg<true>& sa, long long slot, bool valid, bool is_new) {
  // a key created by this launch sits in a slot whose score is 0 (free slots always carry score 0)
  if (valid && slot >= 0) sa.sc[slot] = rule_score(sa.rule, is_new ? 0ull : sa.sc[slot], false, 0ull, now_ns());
}

template <int OPT, bool SCORED = false>
__global__ void __launch_bounds__(kThreadsF)
apply_staged_kernel(TableView t, const long long* __restrict__ keys, const float* __restrict__ grads, size_t n,
                    OptHyper h, const float* __restrict__ init_param, int full_init, unsigned vpr, unsigned lpr,
                    unsigned lpr_shift, int use_tma, ScoreArg<SCORED> sa, const long long* __restrict__ n_dev) {
  if (n_dev) n = (size_t)*n_dev;        // key count produced on the device (det_apply_*_dup); the grid covers the bound
  constexpr int NS = OPT == 0 ? 3 : 4;  // streams per row: grad, param, slot1 (, slot2)
  DET_DYN_SHARED(dyn_smem);
  __shared__ __align__(128) long long s_keys[kStages][kTileKeys];
  __shared__ __align__(8) unsigned long long s_bar[kStages];
  __shared__ unsigned s_new, s_used;
  if (threadIdx.x == 0) {
    s_new = 0;
    s_used = 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned dim = t.dim;
  const unsigned rows_per_step = 32u >> lpr_shift;
  const unsigned sub = (unsigned)lane >> lpr_shift;
  const unsigned c0 = (unsigned)lane & (lpr - 1u);
  const bool lane_on = c0 < vpr;
  float* P = (float*)t.planes[0];
  float* S1 = (float*)t.planes[1];
  float* S2 = (float*)t.planes[2];
  const float omb1 = 1.f - h.beta1, omb2 = 1.f - h.beta2;
  // per warp: 2 stages x kStageSteps steps x NS streams x 32 lanes x 16 B
  float4* wbuf = reinterpret_cast<float4*>(dyn_smem) + (size_t)warp * (2 * kStageSteps * NS * 32);
  auto sm = [&](int stage, int step, int stream) -> float4* { return wbuf + ((stage * kStageSteps + step) * NS + stream) * 32 + lane; };
  // first word of the slot row of MY row-group (lane sub*lpr holds vector 0)
  auto sm_mark = [&](int stage, int step) -> const float* {
    return reinterpret_cast<const float*>(wbuf + ((stage * kStageSteps + step) * NS + 2) * 32 + (sub << lpr_shift));
  };
  const unsigned steps_total = 32u / rows_per_step;            // row-steps per 32 keys
  const unsigned n_batches = (steps_total + kStageSteps - 1) / kStageSteps;
  KeyTiles kt;
  kt.init(s_keys, s_bar, keys, n, use_tma != 0);
  for (; kt.valid(); kt.next()) {
    size_t i;
    bool valid;
    const long long key = kt.key(i, valid);
    const size_t base = i - (size_t)lane;
    bool is_new, from_empty;
    const long long slot = warp_find_or_claim(t, key, valid, valid, lane, is_new, from_empty);
    const unsigned bn = __ballot_sync(kFull, is_new), bu = __ballot_sync(kFull, from_empty);
    if (lane == 0 && bn) {
      atomicAdd(&s_new, __popc(bn));
      atomicAdd(&s_used, __popc(bu));
    }
    score_touch<SCORED>(sa, slot, valid, is_new);
    // issue the async loads of one batch of row-steps into `stage`
    auto issue = [&](unsigned b, int stage) {
#pragma unroll
      for (int st = 0; st < kStageSteps; ++st) {
        const unsigned step = b * kStageSteps + st;
        const unsigned j = step * rows_per_step + sub;
        const long long sl = shfl_ll(slot, (int)(j & 31u));
        const bool nw = (bn >> (j & 31u)) & 1u;
        const size_t gi = base + (j & 31u);
        const bool ok = step < steps_total && gi < n && sl >= 0 && lane_on;
        if (ok) {
          const size_t ro = (size_t)sl * dim + (size_t)c0 * 4;
          cp_async16(sm(stage, st, 0), grads + gi * dim + (size_t)c0 * 4);
          cp_async16(sm(stage, st, 1), nw ? (full_init ? init_param + gi * dim : init_param) + (size_t)c0 * 4 : P + ro);
          if (!nw) {
            cp_async16(sm(stage, st, 2), S1 + ro);
            if (OPT == 1) cp_async16(sm(stage, st, 3), S2 + ro);
          }
        }
      }
      cp_async_commit();
    };
    issue(0, 0);
    for (unsigned b = 0; b < n_batches; ++b) {
      const int stage = (int)(b & 1u);
      if (b + 1 < n_batches) {
        issue(b + 1, stage ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
#pragma unroll
      for (int st = 0; st < kStageSteps; ++st) {
        const unsigned step = b * kStageSteps + st;
        const unsigned j = step * rows_per_step + sub;
        const long long sl = shfl_ll(slot, (int)(j & 31u));
        const bool nw = (bn >> (j & 31u)) & 1u;
        const size_t gi = base + (j & 31u);
        const bool row_ok = step < steps_total && gi < n && sl >= 0;
        // slot state absent: key created in this launch, or created by insert/accum and never stepped
        const bool fresh = nw || (row_ok && __float_as_uint(*sm_mark(stage, st)) == kSlotUninit);
        if (row_ok && lane_on) {
          const size_t ro = (size_t)sl * dim + (size_t)c0 * 4;
          FVec<4> g, p, a, b2;
          g.v = *sm(stage, st, 0);
          p.v = *sm(stage, st, 1);
          if (OPT == 0) {
            if (fresh) a.fill(h.init_slot); else a.v = *sm(stage, st, 2);
            a.zip(g, [](float& av, float gv) { av = av + gv * gv; });
            FVec<4> upd = g;
            upd.zip(a, [&h](float& x, float av) { x = (h.lr * x) / (sqrtf(av) + h.eps); });
            p.zip(upd, [](float& pv, float x) { pv = pv - x; });
            a.store(S1 + ro);
            p.store(P + ro);
          } else {
            if (fresh) { a.zero(); b2.zero(); } else { a.v = *sm(stage, st, 2); b2.v = *sm(stage, st, 3); }
            a.zip(g, [omb1](float& mv, float gv) { mv = mv + (gv - mv) * omb1; });
            b2.zip(g, [omb2](float& vv, float gv) { vv = vv + (gv * gv - vv) * omb2; });
            FVec<4> upd = a;
            upd.zip(b2, [&h](float& x, float vv) { x = (x * h.lr) / (sqrtf(vv) + h.eps); });
            p.zip(upd, [](float& pv, float x) { pv = pv - x; });
            a.store(S1 + ro);
            b2.store(S2 + ro);
            p.store(P + ro);
          }
        }
      }
      __syncwarp();  // every lane is done with this stage before the next issue overwrites it
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_new) {
    atomicAdd(&t.st->size, (unsigned long long)s_new);
    atomicAdd(&t.st->used, (unsigned long long)s_used);
  }
}

// ------------------------------------------------------------------------------------------------
// K8: key-hash partition (stable) + row gather/scatter
// ------------------------------------------------------------------------------------------------
constexpr int kPartBlock = 1024;
constexpr int kMaxShards = 64;

__device__ __forceinline__ int owner_of(long long key, int S, int gpu_mode) {
  if (gpu_mode) {
    const int k32 = (int)(key & 0x7fffffffLL);
    return k32 % S;
  }
  long long m = key % (long long)S;  // floor-mod like tf.math.mod
  if (m < 0) m += S;
  return (int)m;
}

__global__ void __launch_bounds__(kPartBlock)
partition_hist_kernel(const long long* __restrict__ keys, size_t n, int S, int gpu_mode,
                      unsigned* __restrict__ hist /*[S][nblocks]*/, size_t nblocks) {
  __shared__ unsigned s_h[kMaxShards];
  if (threadIdx.x < kMaxShards) s_h[threadIdx.x] = 0;
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * kPartBlock + threadIdx.x;
  if (i < n) atomicAdd(&s_h[owner_of(keys[i], S, gpu_mode)], 1u);
  __syncthreads();
  if ((int)threadIdx.x < S) hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_h[threadIdx.x];
}

// one block: for shard-major flattened hist [S*nblocks], exclusive scan -> global start of each
// (shard, block) run; counts per shard
__global__ void __launch_bounds__(kScanBlock)
partition_scan_kernel(unsigned* __restrict__ hist, size_t nblocks, int S, long long* __restrict__ counts_out) {
  __shared__ unsigned s_warp[33];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const size_t total_n = (size_t)S * nblocks;
  for (size_t base = 0; base < total_n; base += kScanBlock) {
    const size_t i = base + threadIdx.x;
    const unsigned v = i < total_n ? hist[i] : 0u;
    unsigned total;
    const unsigned excl = block_exclusive_scan(v, s_warp, total);
    const unsigned long long carry = s_carry;
    if (i < total_n) hist[i] = (unsigned)(carry + excl);
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + total;
    __syncthreads();
  }
  // counts: start of next shard - start of this shard
  const unsigned long long n_all = s_carry;
  if ((int)threadIdx.x < S) {
    const unsigned long long a = hist[(size_t)threadIdx.x * nblocks];
    const unsigned long long b = ((int)threadIdx.x + 1 < S) ? hist[(size_t)(threadIdx.x + 1) * nblocks] : n_all;
    counts_out[threadIdx.x] = (long long)(b - a);
  }
}

__global__ void __launch_bounds__(kPartBlock)
partition_write_kernel(const long long* __restrict__ keys, size_t n, int S, int gpu_mode,
                       const unsigned* __restrict__ hist, size_t nblocks, long long* __restrict__ keys_out,
                       int* __restrict__ perm_out) {
  __shared__ unsigned s_cnt[kPartBlock / 32][kMaxShards];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int q = threadIdx.x; q < (kPartBlock / 32) * kMaxShards; q += kPartBlock) (&s_cnt[0][0])[q] = 0;
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * kPartBlock + threadIdx.x;
  const bool valid = i < n;
  const long long key = valid ? keys[i] : 0;
  const int own = valid ? owner_of(key, S, gpu_mode) : -1;
  const unsigned peers = __match_any_sync(kFull, own);
  const unsigned rank_in_warp = __popc(peers & ((1u << lane) - 1u));
  if (valid && rank_in_warp == 0) s_cnt[w][own] = __popc(peers);
  __syncthreads();
  // exclusive prefix over warps, per shard (S columns, 32 rows): thread s handles shard s
  if ((int)threadIdx.x < S) {
    unsigned run = 0;
    for (int ww = 0; ww < kPartBlock / 32; ++ww) {
      const unsigned c = s_cnt[ww][threadIdx.x];
      s_cnt[ww][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
  if (valid) {
    const size_t dest = (size_t)hist[(size_t)own * nblocks + blockIdx.x] + s_cnt[w][own] + rank_in_warp;
    keys_out[dest] = key;
    perm_out[dest] = (int)i;
  }
}

template <int VEC, bool SCATTER>
__global__ void __launch_bounds__(kThreadsF)
permute_rows_kernel(const unsigned char* __restrict__ in, const int* __restrict__ perm, size_t n,
                    unsigned char* __restrict__ out, RowGeom g) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = ((size_t)blockIdx.x * kThreadsF + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * kThreadsF) >> 5;
  for (size_t base = warp0 * 32; base < n; base += nwarps * 32) {
    const size_t j = base + lane;
    const unsigned char* src = nullptr;
    unsigned char* dst = nullptr;
    if (j < n) {
      const size_t p = (size_t)perm[j];
      src = in + (SCATTER ? j : p) * g.row_bytes;
      dst = out + (SCATTER ? p : j) * g.row_bytes;
    }
    warp_move_rows<VEC>(g, src, dst, lane);
  }
}

static void fgeom(unsigned dim, bool vec4, unsigned min_lpr, unsigned* vpr, unsigned* lpr, unsigned* sh) {
  *vpr = vec4 ? dim / 4 : dim;
  unsigned l = 1, s = 0;
  while ((l < *vpr || l < min_lpr) && l < 32u) {
    l <<= 1;
    ++s;
  }
  *lpr = l;
  *sh = s;
}


// ------------------------------------------------------------------------------------------------
// K9: det_segment_reduce -- out[g] = sum of the rows whose index is g, added IN POSITION ORDER
// ------------------------------------------------------------------------------------------------
// The gradient dedupe of the sparse optimizer path: TF's _deduplicate_indexed_slices = unique +
// unsorted_segment_sum(values, idx, n_unique) (python/ops/dynamic_embedding_optimizer.py:150,184 ->
// _resource_apply_sparse_duplicate_indices; python/ops/data_flow_grad.py:65; python/ops/math_grad.py:30), whose CPU
// kernel adds the rows of one output row in increasing position -- the order the oracle restates (np.add.at).
// Atomics would make the fp32 sum depend on the schedule, so the rows are GROUPED first:
//   1. stable LSD radix sort of the positions by index, 8 bits per pass over the bits of n_groups only
//      (per pass: radix_hist -> radix_binscan -> radix_scatter; ranks inside a tile come from __match_any_sync,
//      so equal indices keep their position order),
//   2. group_starts: first sorted position of every group;  group_long: the groups longer than kLongGroup,
//   3. segment_reduce_kernel: CTAs [0, n_long_ctas) take the long groups (Zipf head: thousands of rows of one key;
//      the whole CTA stages a tile of rows in shared memory with every load in flight, then thread c adds column c
//      in order), all other CTAs give each lane-group two short groups at a time, 4 row loads in flight per group.
// Indices outside [0, n_groups) are dropped like unsorted_segment_sum drops negative ids.
// 8-bit digits, measured: the c3 sort (1.7M items, 19 significant bits) takes 3 x (9 + 10 + 35) us = 162 us; with 11-bit
// digits (-DDET_RADIX_BITS=11: 2 passes) 2 x (19 + 10 + 60) = 178 us -- the 2048-bin scatter pays for its shared-memory
// counters (profiles/r02_c3_launch_list_{before,after}.csv).  The kernels below are written for any width <= 11.
#ifndef DET_RADIX_BITS
#define DET_RADIX_BITS 8
#endif
constexpr int kRadixBits = DET_RADIX_BITS;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kRadixThreads = 256;
constexpr int kRadixBpt = kRadixBins / kRadixThreads;   // bins per thread: thread t owns bins [t * kRadixBpt, (t + 1) * kRadixBpt)
constexpr int kRadixItems = 8;                     // items per thread
constexpr int kRadixTile = kRadixThreads * kRadixItems;
constexpr int kLongGroup = 64;                     // groups with more rows go to the CTA-cooperative path
constexpr int kHugeGroup = 1024;                   // ... and beyond this a group is split into 16-column slices (one work item each)
constexpr int kSliceCols = 16;                     // 64 B of every row: still two full 32 B sectors per access
constexpr int kLongTileFloats = 8192;              // 32 KB of staged rows per CTA
constexpr int kLongTileRows = 256;                 // staged positions per tile (<= kLongTileFloats / columns)
constexpr int kLongCols = 1024;                    // columns staged at a time (rows wider than this: several sweeps)

// key of position i in pass 0: the index itself, or n_groups ("dropped") when it lies outside [0, n_groups)
__device__ __forceinline__ unsigned radix_key0(const int* __restrict__ idx, size_t i, unsigned n_groups) {
  const unsigned k = (unsigned)idx[i];
  return k < n_groups ? k : n_groups;
}

// hist[bin * nblocks + block] = items of this block's tile whose digit is `bin`
__global__ void __launch_bounds__(kRadixThreads)
radix_hist_kernel(const int* __restrict__ idx, const unsigned* __restrict__ keys_in, size_t n, unsigned n_groups,
                  int shift, unsigned* __restrict__ hist, size_t nblocks, const long long* __restrict__ n_dev) {
  if (n_dev) n = (size_t)*n_dev;
  __shared__ unsigned s_h[kRadixBins];
#pragma unroll
  for (int q = 0; q < kRadixBpt; ++q) s_h[threadIdx.x + q * kRadixThreads] = 0;
  __syncthreads();
  const size_t tile0 = (size_t)blockIdx.x * kRadixTile;
#pragma unroll
  for (int it = 0; it < kRadixItems; ++it) {
    const size_t i = tile0 + (size_t)it * kRadixThreads + threadIdx.x;
    if (i < n) {
      const unsigned k = keys_in ? keys_in[i] : radix_key0(idx, i, n_groups);
      atomicAdd(&s_h[(k >> shift) & (kRadixBins - 1)], 1u);
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < kRadixBpt; ++q) {
    const unsigned bin = threadIdx.x + q * kRadixThreads;
    hist[(size_t)bin * nblocks + blockIdx.x] = s_h[bin];
  }
}

// one warp per bin: exclusive scan of the bin's per-block counts in place, bin total -> totals[bin]
__global__ void __launch_bounds__(kRadixThreads)
radix_binscan_kernel(unsigned* __restrict__ hist, size_t nblocks, unsigned* __restrict__ totals) {
  const int lane = threadIdx.x & 31;
  const unsigned bin = (unsigned)((blockIdx.x * kRadixThreads + threadIdx.x) >> 5);
  if (bin >= (unsigned)kRadixBins) return;   // whole warps only (kRadixThreads is a multiple of 32)
  unsigned* h = hist + (size_t)bin * nblocks;
  unsigned carry = 0;
  for (size_t base = 0; base < nblocks; base += 32) {
    const size_t j = base + lane;
    const unsigned v = j < nblocks ? h[j] : 0u;
    unsigned x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned y = __shfl_up_sync(kFull, x, o);
      if (lane >= o) x += y;
    }
    if (j < nblocks) h[j] = carry + x - v;
    carry += __shfl_sync(kFull, x, 31);
  }
  if (lane == 0) totals[bin] = carry;
}

// stable scatter of one tile: destination = start of the bin + start of this block inside the bin + items of the
// same digit that precede the item inside the block
__global__ void __launch_bounds__(kRadixThreads)
radix_scatter_kernel(const int* __restrict__ idx, const unsigned* __restrict__ keys_in,
                     const unsigned* __restrict__ pos_in, size_t n, unsigned n_groups, int shift,
                     const unsigned* __restrict__ hist, size_t nblocks, const unsigned* __restrict__ totals,
                     unsigned* __restrict__ keys_out, unsigned* __restrict__ pos_out, const long long* __restrict__ n_dev) {
  if (n_dev) n = (size_t)*n_dev;
  constexpr int kWarps = kRadixThreads / 32;
  // per warp: items of each digit (then: exclusive over warps); a warp's sub-tile has 256 items, 16 bits are enough
  __shared__ unsigned short s_cnt[kWarps][kRadixBins];
  __shared__ unsigned s_base[kRadixBins];          // global start of the (bin, block) run
  __shared__ unsigned s_wsum[kWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int q = threadIdx.x; q < kWarps * kRadixBins; q += kRadixThreads) (&s_cnt[0][0])[q] = 0;
  // exclusive scan of the bin totals over all bins: thread t owns kRadixBpt consecutive bins
  unsigned tot[kRadixBpt], mine = 0;
#pragma unroll
  for (int q = 0; q < kRadixBpt; ++q) {
    tot[q] = totals[threadIdx.x * kRadixBpt + q];
    mine += tot[q];
  }
  unsigned x = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned y = __shfl_up_sync(kFull, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_wsum[w] = x;
  __syncthreads();
  unsigned before = 0;
  for (int ww = 0; ww < w; ++ww) before += s_wsum[ww];
  {
    unsigned run = before + x - mine;              // items in the bins before this thread's first bin
#pragma unroll
    for (int q = 0; q < kRadixBpt; ++q) {
      const unsigned bin = threadIdx.x * kRadixBpt + q;
      s_base[bin] = run + hist[(size_t)bin * nblocks + blockIdx.x];
      run += tot[q];
    }
  }
  // pass A: the warp walks its 256 consecutive items 32 at a time; rank inside the warp's sub-tile
  const size_t sub0 = (size_t)blockIdx.x * kRadixTile + (size_t)w * (kRadixTile / kWarps);
  unsigned key[kRadixItems], rank[kRadixItems];
#pragma unroll
  for (int it = 0; it < kRadixItems; ++it) {
    const size_t i = sub0 + (size_t)it * 32 + lane;
    const bool valid = i < n;
    key[it] = valid ? (keys_in ? keys_in[i] : radix_key0(idx, i, n_groups)) : 0u;
    const int d = valid ? (int)((key[it] >> shift) & (kRadixBins - 1)) : -1;
    const unsigned peers = __match_any_sync(kFull, d);
    const unsigned old = valid ? (unsigned)s_cnt[w][d] : 0u;
    rank[it] = old + (unsigned)__popc(peers & ((1u << lane) - 1u));
    __syncwarp();
    if (valid && (peers & ((1u << lane) - 1u)) == 0) s_cnt[w][d] = (unsigned short)(old + (unsigned)__popc(peers));
    __syncwarp();
  }
  __syncthreads();
  {  // exclusive prefix over the warps, per digit (thread t owns digits t, t + 256, ...)
#pragma unroll
    for (int q = 0; q < kRadixBpt; ++q) {
      const unsigned dgt = threadIdx.x + q * kRadixThreads;
      unsigned run = 0;
#pragma unroll
      for (int ww = 0; ww < kWarps; ++ww) {
        const unsigned c = s_cnt[ww][dgt];
        s_cnt[ww][dgt] = (unsigned short)run;
        run += c;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < kRadixItems; ++it) {
    const size_t i = sub0 + (size_t)it * 32 + lane;
    if (i < n) {
      const unsigned d = (key[it] >> shift) & (kRadixBins - 1);
      const size_t dest = (size_t)s_base[d] + s_cnt[w][d] + rank[it];
      keys_out[dest] = key[it];
      pos_out[dest] = pos_in ? pos_in[i] : (unsigned)i;
    }
  }
}

// starts[g] = first sorted position whose key is >= g (g = 0..n_groups); keys are sorted, "dropped" rows carry n_groups
// `ngd` (nullable): the group count lives on the device (det_apply_*_dup: n_unique of det_unique, never read by the host)
__global__ void group_starts_kernel(const unsigned* __restrict__ keys, size_t n, unsigned n_groups,
                                    unsigned* __restrict__ starts, const long long* __restrict__ ngd,
                                    const long long* __restrict__ n_dev, unsigned* __restrict__ ctr) {
  if (ngd) n_groups = (unsigned)*ngd;
  if (n_dev) n = (size_t)*n_dev;
  if (blockIdx.x == 0 && threadIdx.x < 4) ctr[threadIdx.x] = 0u;   // work-list counters of the next two kernels (no memset launch)
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > n) return;
  const long long cur = j < n ? (long long)keys[j] : (long long)n_groups;
  const long long prev = j > 0 ? (long long)keys[j - 1] : -1;
  for (long long g = prev + 1; g <= cur; ++g) starts[g] = (unsigned)j;
}

// work lists of the CTA-cooperative path: ctr[0] = long groups (kLongGroup < rows), ctr[1] = huge groups (> kHugeGroup rows
// of sliceable rows), ctr[2] = the work queue's cursor
__global__ void group_long_kernel(const unsigned* __restrict__ starts, unsigned n_groups, int sliceable,
                                  unsigned* __restrict__ long_list, unsigned* __restrict__ huge_list,
                                  unsigned* __restrict__ ctr, const long long* __restrict__ ngd) {
  if (ngd) n_groups = (unsigned)*ngd;
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const unsigned len = starts[g + 1] - starts[g];
  if (sliceable && len > (unsigned)kHugeGroup) huge_list[atomicAdd(&ctr[1], 1u)] = (unsigned)g;
  else if (len > (unsigned)kLongGroup) long_list[atomicAdd(&ctr[0], 1u)] = (unsigned)g;
}

template <int VF>
__global__ void __launch_bounds__(kThreadsF)
segment_reduce_kernel(const float* __restrict__ rows, const unsigned* __restrict__ pos,
                      const unsigned* __restrict__ starts, unsigned n_groups, unsigned dim, unsigned vpr, unsigned lpr,
                      unsigned lpr_shift, const unsigned* __restrict__ long_list, const unsigned* __restrict__ huge_list,
                      unsigned* __restrict__ ctr, int n_long_ctas, float* __restrict__ out,
                      const long long* __restrict__ ngd) {
  if (ngd) n_groups = (unsigned)*ngd;
  if ((int)blockIdx.x < n_long_ctas) {
    // ---- long groups: a CTA per work item, rows staged tile by tile, thread c adds column c in position order ----
    // Work items come from a queue (atomic cursor), the HUGE groups first: the Zipf head owns thousands of rows of
    // one key, and a sum in position order is a serial chain per column (>= rows x 4 cycles), so the longest group is
    // the critical path of the launch.  A huge group is therefore cut into 16-column slices, one work item each: every
    // slice streams 256 rows per staged tile instead of 64 and the slices of one group run on different SMs.
    __shared__ __align__(16) float s_rows[kLongTileFloats];
    __shared__ unsigned s_pos[kLongTileRows];
    __shared__ unsigned s_q;
    const unsigned n_long = ctr[0], n_huge = ctr[1];
    const unsigned slices = dim / (unsigned)kSliceCols;          // only used when huge groups exist (sliceable rows)
    const unsigned n_huge_items = n_huge * slices;
    while (true) {
      __syncthreads();
      if (threadIdx.x == 0) s_q = atomicAdd(&ctr[2], 1u);
      __syncthreads();
      const unsigned q = s_q;
      if (q >= n_huge_items + n_long) break;
      unsigned g, c_lo, c_hi;
      if (q < n_huge_items) {
        g = huge_list[q / slices];
        c_lo = (q % slices) * (unsigned)kSliceCols;
        c_hi = c_lo + (unsigned)kSliceCols;
      } else {
        g = long_list[q - n_huge_items];
        c_lo = 0;
        c_hi = dim;
      }
      const unsigned st = starts[g], en = starts[g + 1];
      for (unsigned c0 = c_lo; c0 < c_hi; c0 += kLongCols) {
        const unsigned wcols = c_hi - c0 < (unsigned)kLongCols ? c_hi - c0 : (unsigned)kLongCols;  // multiple of VF
        unsigned rpt = (unsigned)kLongTileFloats / wcols;
        if (rpt > (unsigned)kLongTileRows) rpt = kLongTileRows;
        const unsigned wv = wcols / VF;                                   // vectors per staged row
        float acc[kLongCols / kThreadsF];
#pragma unroll
        for (int a = 0; a < kLongCols / kThreadsF; ++a) acc[a] = 0.f;
        if (VF == 4) {
          // double-buffered: tile j+1 is in flight (cp.async, no registers held) while tile j is summed, and the
          // positions of tile j+2 are already being loaded; a tile is half of s_rows (<= 1024 vectors: 4 per thread)
          constexpr unsigned kHalf = kLongTileFloats / 2;
          const unsigned rp2 = kHalf / wcols;                               // rows per tile (>= 4)
          const unsigned ntiles = (en - st + rp2 - 1) / rp2;
          unsigned pp[4];
          auto load_pos = [&](unsigned tile) {
            const unsigned tb = st + tile * rp2;
            const unsigned m = tile < ntiles ? (en - tb < rp2 ? en - tb : rp2) : 0u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const unsigned f = threadIdx.x + (unsigned)i * kThreadsF;
              pp[i] = f < m * wv ? __ldg(pos + tb + f / wv) : 0xffffffffu;
            }
          };
          auto issue = [&](unsigned tile) {
            float* buf = s_rows + (tile & 1u) * kHalf;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const unsigned f = threadIdx.x + (unsigned)i * kThreadsF;
              const unsigned r = f / wv, cv = f - r * wv;
              if (pp[i] != 0xffffffffu)
                cp_async16(buf + (size_t)r * wcols + (size_t)cv * 4, rows + (size_t)pp[i] * dim + c0 + (size_t)cv * 4);
            }
            cp_async_commit();
          };
          load_pos(0);
          issue(0);
          load_pos(1);
          for (unsigned j = 0; j < ntiles; ++j) {
            issue(j + 1);                  // an empty group past the last tile keeps the wait count uniform
            load_pos(j + 2);
            cp_async_wait<1>();
            __syncthreads();
            const unsigned tb = st + j * rp2;
            const unsigned m = en - tb < rp2 ? en - tb : rp2;
            const float* buf = s_rows + (j & 1u) * kHalf;
#pragma unroll
            for (int a = 0; a < kLongCols / kThreadsF; ++a) {
              const unsigned c = threadIdx.x + (unsigned)a * kThreadsF;
              if (c < wcols) {
                float sacc = acc[a];
                for (unsigned r = 0; r < m; ++r) sacc = sacc + buf[(size_t)r * wcols + c];
                acc[a] = sacc;
              }
            }
            __syncthreads();               // tile j's buffer is free for tile j+2
          }
          cp_async_wait<0>();
        } else {
          for (unsigned base = st; base < en; base += rpt) {
            const unsigned m = en - base < rpt ? en - base : rpt;
            for (unsigned r = threadIdx.x; r < m; r += kThreadsF) s_pos[r] = pos[base + r];
            __syncthreads();
            const unsigned total = m * wv;
  #pragma unroll 4
            for (unsigned f = threadIdx.x; f < total; f += kThreadsF) {
              const unsigned r = f / wv, cv = f - r * wv;
              FVec<VF> x;
              x.load(rows + (size_t)s_pos[r] * dim + c0 + (size_t)cv * VF);
              x.store(s_rows + (size_t)r * wcols + (size_t)cv * VF);
            }
            __syncthreads();
  #pragma unroll
            for (int a = 0; a < kLongCols / kThreadsF; ++a) {
              const unsigned c = threadIdx.x + (unsigned)a * kThreadsF;
              if (c < wcols) {
                float s = acc[a];
                for (unsigned r = 0; r < m; ++r) s = s + s_rows[(size_t)r * wcols + c];
                acc[a] = s;
              }
            }
            __syncthreads();
          }
        }
#pragma unroll
        for (int a = 0; a < kLongCols / kThreadsF; ++a) {
          const unsigned c = threadIdx.x + (unsigned)a * kThreadsF;
          if (c < wcols) out[(size_t)g * dim + c0 + c] = acc[a];
        }
      }
    }
    return;
  }
  // ---- short groups: a lane-group walks U groups in lock step, K rows of each in flight ----
  constexpr int U = 2, K = 4;
  const int lane = threadIdx.x & 31;
  const unsigned gl = (unsigned)lane & (lpr - 1u);
  const unsigned gpw = 32u >> lpr_shift;
  const unsigned gidx = (unsigned)lane >> lpr_shift;
  const size_t warp0 = (((size_t)blockIdx.x - (size_t)n_long_ctas) * kThreadsF + threadIdx.x) >> 5;
  const size_t nwarps = (((size_t)gridDim.x - (size_t)n_long_ctas) * kThreadsF) >> 5;
  const unsigned nchunk = (vpr + lpr - 1u) / lpr;   // 1 unless the row is wider than 32 vectors
  // the U + 1 boundaries of the NEXT iteration's groups are loaded one iteration ahead (they head a chain of three
  // dependent loads: boundary -> position -> row)
  unsigned nb[U + 1];
  auto load_bounds = [&](size_t sb_) {
    const size_t g0_ = sb_ + (size_t)gidx * U;
#pragma unroll
    for (int u = 0; u <= U; ++u) nb[u] = (sb_ < n_groups && g0_ + u <= n_groups) ? __ldg(starts + g0_ + u) : 0u;
  };
  load_bounds(warp0 * gpw * U);
  for (size_t sb = warp0 * gpw * U; sb < n_groups; sb += nwarps * gpw * U) {
    const size_t g0 = sb + (size_t)gidx * U;
    unsigned st[U], en[U];
    bool mine[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = g0 + u < n_groups;
      const unsigned a = ok ? nb[u] : 0u, b = ok ? nb[u + 1] : 0u;
      mine[u] = ok && b - a <= (unsigned)kLongGroup;
      st[u] = mine[u] ? a : 0u;
      en[u] = mine[u] ? b : 0u;
    }
    load_bounds(sb + nwarps * gpw * U);
    for (unsigned cc = 0; cc < nchunk; ++cc) {
      const unsigned cv = cc * lpr + gl;
      const bool lane_on = cv < vpr;
      FVec<VF> acc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) acc[u].zero();
      for (unsigned k = 0;; k += K) {
        bool any = false;
#pragma unroll
        for (int u = 0; u < U; ++u) any |= st[u] + k < en[u];
        if (!__any_sync(kFull, any)) break;
        unsigned p[U][K];
        FVec<VF> x[U][K];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int q = 0; q < K; ++q) p[u][q] = st[u] + k + q < en[u] ? __ldg(pos + st[u] + k + q) : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int q = 0; q < K; ++q) {
            x[u][q].zero();
            if (p[u][q] != 0xffffffffu && lane_on) x[u][q].load(rows + (size_t)p[u][q] * dim + (size_t)cv * VF);
          }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int q = 0; q < K; ++q)
            if (p[u][q] != 0xffffffffu) acc[u].zip(x[u][q], [](float& a, float xv) { a = a + xv; });
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (mine[u] && lane_on) acc[u].store(out + (g0 + u) * dim + (size_t)cv * VF);
    }
  }
}

struct SegReduceWs {
  unsigned *keys_a, *keys_b, *pos_a, *pos_b, *hist, *totals, *starts, *long_list, *huge_list, *n_long;
};

static size_t seg_reduce_layout(size_t n, size_t n_groups, unsigned char* base, SegReduceWs* w) {
  const size_t nblocks = (n + kRadixTile - 1) / kRadixTile;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off += align256(bytes);
    return (unsigned*)p;
  };
  unsigned* ka = take(n * 4);
  unsigned* kb = take(n * 4);
  unsigned* pa = take(n * 4);
  unsigned* pb = take(n * 4);
  unsigned* hist = take((size_t)kRadixBins * nblocks * 4);
  unsigned* totals = take(kRadixBins * 4);
  unsigned* starts = take((n_groups + 2) * 4);
  unsigned* ll = take((n / kLongGroup + 2) * 4);
  unsigned* hl = take((n / kHugeGroup + 2) * 4);
  unsigned* nl = take(16);
  if (w) {
    w->keys_a = ka; w->keys_b = kb; w->pos_a = pa; w->pos_b = pb; w->hist = hist; w->totals = totals;
    w->starts = starts; w->long_list = ll; w->huge_list = hl; w->n_long = nl;
  }
  return off;
}

}  // namespace det

using namespace det;

extern "C" {

size_t det_unique_workspace_bytes(size_t n) { return unique_ws_layout(n ? n : 1, nullptr, nullptr); }

// n bounds the id count (sizes, grids); with `n_dev` the real count is read on the device by every kernel
static det_status unique_run(const int64_t* ids, size_t n, int64_t* unique_out, int32_t* idx_out, int64_t* n_unique_dev,
                             void* workspace, size_t workspace_bytes, const long long* n_dev, cudaStream_t s) {
  if (!n_unique_dev) return fail(DET_INVALID_ARGUMENT, "det_unique: null n_unique");
  if (n == 0) {
    CUDA_TRY(cudaMemsetAsync(n_unique_dev, 0, sizeof(int64_t), s));
    return DET_OK;
  }
  if (n >= 0xfffffff0ull) return fail(DET_INVALID_ARGUMENT, "det_unique: n too large");
  if (!ids || !unique_out || !idx_out || !workspace) return fail(DET_INVALID_ARGUMENT, "det_unique: null argument");
  if (workspace_bytes < det_unique_workspace_bytes(n)) return fail(DET_INVALID_ARGUMENT, "det_unique: workspace too small");
  UniqueWs w;
  unique_ws_layout(n, (unsigned char*)workspace, &w);
  const size_t nblocks = (n + kScanBlock - 1) / kScanBlock;
  const int g256 = (int)((n + 255) / 256);
  DET_LAUNCH(unique_init_kernel, (int)((w.hcap + 1023) / 1024 < 4096 ? (w.hcap + 1023) / 1024 : 4096), 1024, 0, s, w);
  DET_LAUNCH(unique_insert_kernel, g256, 256, 0, s, w, (const long long*)ids, n, n_dev);
  DET_LAUNCH(unique_flag_sums_kernel, (int)nblocks, kScanBlock, 0, s, w, n, n_dev);
  DET_LAUNCH(scan_sums_kernel, 1, kScanBlock, 0, s, w.block_sums, nblocks, (long long*)n_unique_dev);
  DET_LAUNCH(unique_rank_kernel, (int)nblocks, kScanBlock, 0, s, w, (const long long*)ids, n, (long long*)unique_out, n_dev);
  DET_LAUNCH(unique_idx_kernel, g256, 256, 0, s, w, n, idx_out, n_dev);
  CUDA_TRY(cudaGetLastError());
  return DET_OK;
}

det_status det_unique(const int64_t* ids, size_t n, int64_t* unique_out, int32_t* idx_out, int64_t* n_unique_dev,
                      void* workspace, size_t workspace_bytes, det_stream_t stream) {
  return unique_run(ids, n, unique_out, idx_out, n_unique_dev, workspace, workspace_bytes, nullptr, (cudaStream_t)stream);
}

static det_status lookup_sparse_impl(det_table* t, const int64_t* ids, const int32_t* segment_ids, const float* weights,
                                     size_t nnz, size_t batch, int combiner, const float* default_row, float max_norm,
                                     float* out, det_stream_t stream) {
  if (!t) return fail(DET_INVALID_ARGUMENT, "det_lookup_sparse: null table");
  if (!(max_norm >= 0.f)) return fail(DET_INVALID_ARGUMENT, "det_lookup_sparse: max_norm must be >= 0 (0 = no clipping)");
  std::lock_guard<std::mutex> _lk(t->mu);
  if (t->cfg.value_dtype != DET_FLOAT32) return fail(DET_UNIMPLEMENTED, "det_lookup_sparse: float32 tables only");
  if (combiner < DET_COMBINER_SUM || combiner > DET_COMBINER_SQRTN)
    return fail(DET_INVALID_ARGUMENT, "combiner must be one of 'mean', 'sqrtn' or 'sum'");
  if (batch == 0) return DET_OK;
  if (!out || !default_row || (nnz && (!ids || !segment_ids))) return fail(DET_INVALID_ARGUMENT, "det_lookup_sparse: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  det::DevGuard _dg(t->cfg.device);
  long long* seg_start = nullptr;
  long long* slots = nullptr;
  unsigned* not_identity = nullptr;
  {
    void* sc = nullptr;
    const size_t seg_bytes = ((batch + 1) * sizeof(long long) + 255) & ~(size_t)255;
    det_status sst = table_scratch(t, 256 + seg_bytes + (nnz ? nnz : 1) * sizeof(long long), &sc);
    if (sst != DET_OK) return sst;
    not_identity = (unsigned*)sc;
    seg_start = (long long*)((unsigned char*)sc + 256);
    slots = (long long*)((unsigned char*)sc + 256 + seg_bytes);
  }
  const unsigned dim = (unsigned)t->cfg.dim;
  const bool vec4 = (dim % 4 == 0) && ((((uintptr_t)default_row | (uintptr_t)out) & 15u) == 0);
  unsigned vpr, lpr, sh;
  fgeom(dim, vec4, 1, &vpr, &lpr, &sh);
  // One id per output row and no weights (the Criteo shape: 26 features x batch, one id each): the op is a Find into the
  // dense output.  Whether the segment ids really are 0..nnz-1 is only known on the DEVICE (segment_offsets_kernel
  // computes it): the find-shaped kernel and the general pair are both launched, the flag lets exactly one of them work.
  const int ivec = pick_vec(t->row_bytes, default_row, out, nullptr);
  const bool try_identity = nnz == batch && nnz > 0 && weights == nullptr && !(max_norm > 0.f) && vpr <= lpr && ivec >= 4 &&
                            env_int("DET_SEGSUM_STAGED", 0) == 0 && env_int("DET_SPARSE_IDENTITY", 1) != 0;
  const unsigned* gate = try_identity ? not_identity : nullptr;
  // the flag word is compared with a per-table call counter instead of being reset by a memset launch per call
  unsigned epoch = 0;
  if (try_identity) {
    if (t->sparse_flag != (void*)not_identity) {     // scratch (re)allocated: unknown content, clear it once
      CUDA_TRY(cudaMemsetAsync(not_identity, 0, sizeof(unsigned), s));
      t->sparse_flag = (void*)not_identity;
      t->sparse_epoch = 0;
    }
    epoch = ++t->sparse_epoch;
    if (epoch == 0) epoch = ++t->sparse_epoch;
    const int vec = ivec;
    const RowGeom g = make_geom((unsigned)t->row_bytes, vec);
    const int tma = (((uintptr_t)ids & 15u) == 0) ? 1 : 0;
    const long long* k = (const long long*)ids;
    const unsigned char* d = (const unsigned char*)default_row;
    unsigned char* o = (unsigned char*)out;
#define DET_IDENT(VV)                                                                                              \
  case VV: {                                                                                                        \
    const int grid = grid_for(nnz, kTileKeys, t->sm_count, occupancy_of(lookup_identity_kernel<VV>, kThreadsF));   \
    DET_LAUNCH(lookup_identity_kernel<VV>, grid, kThreadsF, 0, s, t->view, k, segment_ids, nnz, d, o, g, tma, not_identity, epoch); \
  } break;
    switch (vec) {
      DET_IDENT(16) DET_IDENT(8) DET_IDENT(4)
      default: break;
    }
#undef DET_IDENT
  }
  DET_LAUNCH(segment_offsets_kernel, (int)((nnz + 1 + 255) / 256), 256, 0, s, segment_ids, nnz, batch, seg_start, t->view.st, gate, epoch);
  if (nnz)
    DET_LAUNCH(resolve_slots_kernel, grid_for(nnz, kTileKeys, t->sm_count, occupancy_of(resolve_slots_kernel, kThreadsF)), kThreadsF, 0, s, t->view, (const long long*)ids, nnz, slots, (((uintptr_t)ids & 15u) == 0) ? 1 : 0, gate, epoch);
  det_status rc = DET_OK;
  const unsigned gpw = 32u >> sh;
  if (vpr <= lpr && max_norm > 0.f) {
    ClipArg<true> clip;
    clip.max_norm = max_norm;
    const auto k4 = segment_sum_kernel<4, true>;
    const auto k1 = segment_sum_kernel<1, true>;
    const int occ = vec4 ? occupancy_of(k4, kThreadsF) : occupancy_of(k1, kThreadsF);
    const int grid = grid_for(batch, (int)(gpw * kSegPerGroup * (kThreadsF / 32)), t->sm_count, occ);
    if (vec4)
      DET_LAUNCH(k4, grid, kThreadsF, 0, s, t->view, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh, clip, (const unsigned*)nullptr, 0u);
    else
      DET_LAUNCH(k1, grid, kThreadsF, 0, s, t->view, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh, clip, (const unsigned*)nullptr, 0u);
  } else if (max_norm > 0.f) {
    rc = fail(DET_UNIMPLEMENTED, "det_lookup_sparse_clip: max_norm is fused for rows of at most 32 vectors (dim <= 128 when "
                                 "dim % 4 == 0, else dim <= 32); use the composed path for wider rows");
  } else if (vpr <= lpr && vec4 && env_int("DET_SEGSUM_STAGED", 0) != 0) {
    // round-2 candidate (default off): rows staged through shared memory with cp.async, see segment_sum_staged_kernel
    const size_t smem = (size_t)(kThreadsF / 32) * kWinWarpBytes;
    int occ = 1;
    CUDA_TRY(cudaFuncSetAttribute(segment_sum_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, segment_sum_staged_kernel, kThreadsF, smem) != cudaSuccess || occ < 1) occ = 1;
    const int grid = grid_for(batch, kThreadsF / 32, t->sm_count, occ);
#ifdef DET_EMU
    __atomic_fetch_add(&g_det_emu_stat[2], 1ull, __ATOMIC_RELAXED);
#endif
    DET_LAUNCH(segment_sum_staged_kernel, grid, kThreadsF, smem, s, t->view, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh);
  } else if (vpr <= lpr) {
    const ClipArg<false> noclip;
    const int occ = vec4 ? occupancy_of(segment_sum_kernel<4>, kThreadsF) : occupancy_of(segment_sum_kernel<1>, kThreadsF);
    const int grid = grid_for(batch, (int)(gpw * kSegPerGroup * (kThreadsF / 32)), t->sm_count, occ);
    if (vec4)
      DET_LAUNCH(segment_sum_kernel<4>, grid, kThreadsF, 0, s, t->view, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh, noclip, gate, epoch);
    else
      DET_LAUNCH(segment_sum_kernel<1>, grid, kThreadsF, 0, s, t->view, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh, noclip, gate, epoch);
  } else if ((vpr + lpr - 1) / lpr <= (unsigned)kMaxVecPerLane) {
    const int grid = grid_for(batch, (int)(gpw * (kThreadsF / 32)), t->sm_count, 8);
    if (vec4)
      DET_LAUNCH(segment_sum_wide_kernel<4>, grid, kThreadsF, 0, s, t->view, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh);
    else
      DET_LAUNCH(segment_sum_wide_kernel<1>, grid, kThreadsF, 0, s, t->view, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh);
  } else {
    rc = fail(DET_UNIMPLEMENTED, "det_lookup_sparse: dim too large for the fused kernel");
  }
  if (rc == DET_OK && cudaGetLastError() != cudaSuccess) rc = fail(DET_CUDA_ERROR, "det_lookup_sparse: launch failed");
  return rc;
}

det_status det_lookup_sparse(det_table* t, const int64_t* ids, const int32_t* segment_ids, const float* weights,
                             size_t nnz, size_t batch, int combiner, const float* default_row, float* out,
                             det_stream_t stream) {
  return lookup_sparse_impl(t, ids, segment_ids, weights, nnz, batch, combiner, default_row, 0.f, out, stream);
}

det_status det_lookup_sparse_clip(det_table* t, const int64_t* ids, const int32_t* segment_ids, const float* weights,
                                  size_t nnz, size_t batch, int combiner, const float* default_row, float max_norm,
                                  float* out, det_stream_t stream) {
  return lookup_sparse_impl(t, ids, segment_ids, weights, nnz, batch, combiner, default_row, max_norm, out, stream);
}

// n_dev (nullable): the key count lives on the device (det_apply_*_dup); n is then its upper bound -- room is reserved
// for n NEW keys (the unique keys themselves are not read by the host path) and only the staged kernels support it
static det_status apply_common(det_table* t, const int64_t* keys, const float* grads, size_t n, OptHyper h,
                               const float* init_param, int full_init, int opt, cudaStream_t s,
                               const long long* n_dev = nullptr) {
  if (!t) return fail(DET_INVALID_ARGUMENT, "det_apply: null table");
  std::lock_guard<std::mutex> _lk(t->mu);
  if (t->cfg.value_dtype != DET_FLOAT32) return fail(DET_UNIMPLEMENTED, "det_apply: float32 tables only");
  if (t->cfg.num_slot_planes < (opt == 0 ? 1 : 2))
    return fail(DET_INVALID_ARGUMENT, "det_apply: table was created with too few optimizer slot planes");
  if (n == 0) return DET_OK;
  if (!keys || !grads || !init_param) return fail(DET_INVALID_ARGUMENT, "det_apply: null argument");
  det::DevGuard _dg(t->cfg.device);
  const unsigned dim = (unsigned)t->cfg.dim;
  const bool vec4 = (dim % 4 == 0) && ((((uintptr_t)grads | (uintptr_t)init_param) & 15u) == 0);
  unsigned vpr, lpr, sh;
  fgeom(dim, vec4, 1, &vpr, &lpr, &sh);
  if (n_dev && !(vec4 && vpr <= lpr))
    return fail(DET_UNIMPLEMENTED, "det_apply_*_dup: rows of up to 32 aligned 16-byte vectors only (use unique + segment_reduce + apply)");
  det_status st = ensure_room(t, n_dev ? nullptr : (const long long*)keys, n, s);
  if (st != DET_OK) return st;
  static const int ru = env_int("DET_APPLY_RU", 1);  // measured on B200: one row-step per iteration wins (more resident CTAs)
  static const int staged = env_int("DET_APPLY_STAGED", 1);  // measured: +12..18 % over the register-held variant
  const TableView v = t->view;
  const long long* k = (const long long*)keys;
  if ((staged || n_dev) && vec4 && vpr <= lpr) {
    // cp.async-staged variant: 2 stages x kStageSteps x streams x 512 B per warp of dynamic shared memory
    const int ns = opt == 0 ? 3 : 4;
    const size_t smem = (size_t)(kThreadsF / 32) * 2 * kStageSteps * ns * 32 * 16;
    const int tma = (((uintptr_t)keys & 15u) == 0) ? 1 : 0;
    int occ = 1;
    const bool scored = t->ev != nullptr;
#define DET_LAUNCH_STAGED(OPT_, SCORED_, SA_)                                                                       \
  {                                                                                                               \
    const auto kern = apply_staged_kernel<OPT_, SCORED_>;                                                         \
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreadsF, smem) != cudaSuccess || occ < 1) occ = 1; \
    DET_LAUNCH(kern, grid_for(n, kTileKeys, t->sm_count, occ), kThreadsF, smem, s, v, k, grads, n, h, init_param, \
               full_init, vpr, lpr, sh, tma, SA_, n_dev);                                                         \
  }
    if (scored) {
      ScoreArg<true> sa;
      sa.sc = evict_scores(t);
      sa.rule = evict_rule(t);
      if (opt == 0) DET_LAUNCH_STAGED(0, true, sa) else DET_LAUNCH_STAGED(1, true, sa)
    } else {
      const ScoreArg<false> sa;
      if (opt == 0) DET_LAUNCH_STAGED(0, false, sa) else DET_LAUNCH_STAGED(1, false, sa)
    }
#undef DET_LAUNCH_STAGED
    CUDA_TRY(cudaGetLastError());
    note_mutation(t, n, s);
    return DET_OK;   // scores of a table with an eviction strategy were written by the kernel itself
  }
#define DET_LAUNCH_APPLY(VF_, OPT_, RU_)                                                                      \
  {                                                                                                           \
    const int grid = grid_for(n, kTileKeys, t->sm_count, occupancy_of(apply_kernel<VF_, OPT_, RU_>, kThreadsF)); \
    const auto kern = apply_kernel<VF_, OPT_, RU_>;                                                           \
    DET_LAUNCH(kern, grid, kThreadsF, 0, s, v, k, grads, n, h, init_param, full_init, vpr, lpr, sh,                   \
                                    (((uintptr_t)keys & 15u) == 0) ? 1 : 0);                                  \
  }
  if (opt == 0) {
    if (vec4) { if (ru == 2) DET_LAUNCH_APPLY(4, 0, 2) else DET_LAUNCH_APPLY(4, 0, 1) }
    else DET_LAUNCH_APPLY(1, 0, 1)
  } else {
    if (vec4) { if (ru == 2) DET_LAUNCH_APPLY(4, 1, 2) else DET_LAUNCH_APPLY(4, 1, 1) }
    else DET_LAUNCH_APPLY(1, 1, 1)
  }
#undef DET_LAUNCH_APPLY
  CUDA_TRY(cudaGetLastError());
  note_mutation(t, n, s);
  if (t->ev) return evict_touch(t, k, nullptr, n, s);
  return DET_OK;
}

det_status det_apply_adagrad(det_table* t, const int64_t* keys, const float* grads, size_t n, float lr, float epsilon,
                             const float* init_param, int full_size_init, float init_accum, det_stream_t stream) {
  OptHyper h;
  h.lr = lr; h.eps = epsilon; h.beta1 = 0.f; h.beta2 = 0.f; h.init_slot = init_accum;
  if (t) t->slot_init[1] = init_accum;
  return apply_common(t, keys, grads, n, h, init_param, full_size_init, 0, (cudaStream_t)stream);
}

det_status det_apply_adam(det_table* t, const int64_t* keys, const float* grads, size_t n, float alpha, float beta1,
                          float beta2, float epsilon, const float* init_param, int full_size_init,
                          det_stream_t stream) {
  OptHyper h;
  h.lr = alpha; h.eps = epsilon; h.beta1 = beta1; h.beta2 = beta2; h.init_slot = 0.f;
  return apply_common(t, keys, grads, n, h, init_param, full_size_init, 1, (cudaStream_t)stream);
}

size_t det_partition_workspace_bytes(size_t n, int num_shards) {
  const size_t nblocks = (n + kPartBlock - 1) / kPartBlock;
  return align256((size_t)(num_shards > 0 ? num_shards : 1) * (nblocks ? nblocks : 1) * sizeof(unsigned)) + 256;
}

det_status det_partition(const int64_t* keys, size_t n, int num_shards, int gpu_mode, int64_t* keys_out,
                         int32_t* perm_out, int64_t* counts_out, void* workspace, size_t workspace_bytes,
                         det_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (num_shards < 1 || num_shards > kMaxShards) return fail(DET_INVALID_ARGUMENT, "det_partition: num_shards must be in [1,64]");
  if (!counts_out) return fail(DET_INVALID_ARGUMENT, "det_partition: null counts");
  if (n == 0) {
    CUDA_TRY(cudaMemsetAsync(counts_out, 0, sizeof(int64_t) * num_shards, s));
    return DET_OK;
  }
  if (n >= 0x7fffffffull) return fail(DET_INVALID_ARGUMENT, "det_partition: n too large");
  if (!keys || !keys_out || !perm_out || !workspace) return fail(DET_INVALID_ARGUMENT, "det_partition: null argument");
  if (workspace_bytes < det_partition_workspace_bytes(n, num_shards)) return fail(DET_INVALID_ARGUMENT, "det_partition: workspace too small");
  const size_t nblocks = (n + kPartBlock - 1) / kPartBlock;
  unsigned* hist = (unsigned*)workspace;
  DET_LAUNCH(partition_hist_kernel, (int)nblocks, kPartBlock, 0, s, (const long long*)keys, n, num_shards, gpu_mode, hist, nblocks);
  DET_LAUNCH(partition_scan_kernel, 1, kScanBlock, 0, s, hist, nblocks, num_shards, (long long*)counts_out);
  DET_LAUNCH(partition_write_kernel, (int)nblocks, kPartBlock, 0, s, (const long long*)keys, n, num_shards, gpu_mode, hist, nblocks,
                                                            (long long*)keys_out, perm_out);
  CUDA_TRY(cudaGetLastError());
  return DET_OK;
}

static det_status permute_rows(const void* in, const int32_t* perm, size_t n, size_t row_bytes, void* out, bool scatter,
                               cudaStream_t s) {
  if (n == 0) return DET_OK;
  if (!in || !perm || !out || row_bytes == 0) return fail(DET_INVALID_ARGUMENT, "det_*_rows: null argument");
  const int vec = pick_vec(row_bytes, in, out, nullptr);
  const RowGeom g = make_geom((unsigned)row_bytes, vec);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = grid_for(n, kThreadsF, sms, 8);
  const unsigned char* i = (const unsigned char*)in;
  unsigned char* o = (unsigned char*)out;
#define LAUNCH(V)                                                        \
  {                                                                      \
    const auto ks = permute_rows_kernel<V, true>;                        \
    const auto kg = permute_rows_kernel<V, false>;                       \
    if (scatter) DET_LAUNCH(ks, grid, kThreadsF, 0, s, i, perm, n, o, g);        \
    else DET_LAUNCH(kg, grid, kThreadsF, 0, s, i, perm, n, o, g);                \
  }
  switch (vec) {
    case 16: LAUNCH(16) break;
    case 8: LAUNCH(8) break;
    case 4: LAUNCH(4) break;
    case 2: LAUNCH(2) break;
    default: LAUNCH(1) break;
  }
#undef LAUNCH
  CUDA_TRY(cudaGetLastError());
  return DET_OK;
}

det_status det_scatter_rows(const void* rows_in, const int32_t* perm, size_t n, size_t row_bytes, void* rows_out,
                            det_stream_t stream) {
  return permute_rows(rows_in, perm, n, row_bytes, rows_out, true, (cudaStream_t)stream);
}

det_status det_gather_rows(const void* rows_in, const int32_t* perm, size_t n, size_t row_bytes, void* rows_out,
                           det_stream_t stream) {
  return permute_rows(rows_in, perm, n, row_bytes, rows_out, false, (cudaStream_t)stream);
}

// tf.sparse.segment_{sum,mean,sqrt_n} with weights over a DENSE row matrix: the tail of embedding_lookup_sparse in
// TRAINING, where the unique rows live in the TrainableWrapper's scratch (python/ops/dynamic_embedding_ops.py:247-289:
// gather(embeddings, idx) * weights -> segment_sum -> / sum(w) | / sqrt(sum(w^2))).  The gather, its weighted copy and
// the segment reduction are ONE pass of the kernels det_lookup_sparse uses for its phase B (segment_sum_kernel /
// segment_sum_wide_kernel, or the staged variant under DET_SEGSUM_STAGED=1) over a view whose "table" is the matrix.
size_t det_sparse_segment_sum_workspace_bytes(size_t batch) {
  return align256((batch + 1) * sizeof(long long)) + align256(sizeof(DevState));
}

det_status det_sparse_segment_sum(const float* rows, size_t dim, const int64_t* row_idx, const int32_t* segment_ids,
                                  const float* weights, size_t nnz, size_t batch, int combiner, const float* default_row,
                                  float* out, void* workspace, size_t workspace_bytes, det_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (batch == 0 || dim == 0) return DET_OK;
  if (!out || !workspace || !default_row) return fail(DET_INVALID_ARGUMENT, "det_sparse_segment_sum: null out/workspace/default_row");
  if (nnz && (!rows || !row_idx || !segment_ids)) return fail(DET_INVALID_ARGUMENT, "det_sparse_segment_sum: null argument");
  if (combiner < DET_COMBINER_SUM || combiner > DET_COMBINER_SQRTN) return fail(DET_INVALID_ARGUMENT, "det_sparse_segment_sum: bad combiner");
  if (workspace_bytes < det_sparse_segment_sum_workspace_bytes(batch))
    return fail(DET_INVALID_ARGUMENT, "det_sparse_segment_sum: workspace too small");
  if (nnz >= 0x7fffffffull || dim >= 0x7fffffffull) return fail(DET_INVALID_ARGUMENT, "det_sparse_segment_sum: nnz / dim too large");
  long long* seg_start = (long long*)workspace;
  DevState* st = (DevState*)((unsigned char*)workspace + align256((batch + 1) * sizeof(long long)));
  CUDA_TRY(cudaMemsetAsync(st, 0, sizeof(DevState), s));
  TableView v{};
  v.planes[0] = (unsigned char*)const_cast<float*>(rows);
  v.dim = (unsigned)dim;
  v.row_bytes = (unsigned)dim * 4u;
  v.st = st;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  DET_LAUNCH(segment_offsets_kernel, (int)((nnz + 1 + 255) / 256), 256, 0, s, segment_ids, nnz, batch, seg_start, st, (const unsigned*)nullptr, 0u);
  const long long* slots = (const long long*)row_idx;
  const bool vec4 = (dim % 4 == 0) && ((((uintptr_t)default_row | (uintptr_t)out | (uintptr_t)rows) & 15u) == 0);
  unsigned vpr, lpr, sh;
  fgeom((unsigned)dim, vec4, 1, &vpr, &lpr, &sh);
  const unsigned gpw = 32u >> sh;
  if (vpr <= lpr && vec4 && env_int("DET_SEGSUM_STAGED", 0) != 0) {
    const size_t smem = (size_t)(kThreadsF / 32) * kWinWarpBytes;
    int occ = 1;
    CUDA_TRY(cudaFuncSetAttribute(segment_sum_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, segment_sum_staged_kernel, kThreadsF, smem) != cudaSuccess || occ < 1) occ = 1;
    const int grid = grid_for(batch, kThreadsF / 32, sms, occ);
#ifdef DET_EMU
    __atomic_fetch_add(&g_det_emu_stat[2], 1ull, __ATOMIC_RELAXED);
#endif
    DET_LAUNCH(segment_sum_staged_kernel, grid, kThreadsF, smem, s, v, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh);
  } else if (vpr <= lpr) {
    const ClipArg<false> noclip;
    const int occ = vec4 ? occupancy_of(segment_sum_kernel<4>, kThreadsF) : occupancy_of(segment_sum_kernel<1>, kThreadsF);
    const int grid = grid_for(batch, (int)(gpw * kSegPerGroup * (kThreadsF / 32)), sms, occ);
    if (vec4)
      DET_LAUNCH(segment_sum_kernel<4>, grid, kThreadsF, 0, s, v, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh, noclip, (const unsigned*)nullptr, 0u);
    else
      DET_LAUNCH(segment_sum_kernel<1>, grid, kThreadsF, 0, s, v, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh, noclip, (const unsigned*)nullptr, 0u);
  } else {
    const int grid = grid_for(batch, (int)(gpw * (kThreadsF / 32)), sms, 8);
    if (vec4)
      DET_LAUNCH(segment_sum_wide_kernel<4>, grid, kThreadsF, 0, s, v, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh);
    else
      DET_LAUNCH(segment_sum_wide_kernel<1>, grid, kThreadsF, 0, s, v, slots, seg_start, weights, batch, combiner, default_row, out, vpr, lpr, sh);
  }
  CUDA_TRY(cudaGetLastError());
  return DET_OK;
}

size_t det_segment_reduce_workspace_bytes(size_t n, size_t n_groups) {
  return seg_reduce_layout(n ? n : 1, n_groups ? n_groups : 1, nullptr, nullptr);
}

// n_groups bounds the group count (sizes, pass count, grids); with `ngd` the real count is read on the device
static det_status seg_reduce_run(const float* rows, const int32_t* idx, size_t n, size_t n_groups, size_t dim, float* out,
                                 void* workspace, size_t workspace_bytes, const long long* ngd, cudaStream_t s,
                                 const long long* n_dev = nullptr) {
  if (n_groups == 0 || dim == 0) return DET_OK;
  if (!out) return fail(DET_INVALID_ARGUMENT, "det_segment_reduce: null out");
  if (n >= 0x7fffffffull || n_groups >= 0x7fffffffull || dim >= 0x7fffffffull)
    return fail(DET_INVALID_ARGUMENT, "det_segment_reduce: n / n_groups / dim too large");
  if (n == 0) {
    if (!ngd) CUDA_TRY(cudaMemsetAsync(out, 0, n_groups * dim * sizeof(float), s));
    return DET_OK;
  }
  if (!rows || !idx || !workspace) return fail(DET_INVALID_ARGUMENT, "det_segment_reduce: null argument");
  if (workspace_bytes < det_segment_reduce_workspace_bytes(n, n_groups))
    return fail(DET_INVALID_ARGUMENT, "det_segment_reduce: workspace too small");
  SegReduceWs w;
  seg_reduce_layout(n, n_groups, (unsigned char*)workspace, &w);
  const size_t nblocks = (n + kRadixTile - 1) / kRadixTile;
  const unsigned ng = (unsigned)n_groups;
  int bits = 0;
  for (size_t v = n_groups; v; v >>= 1) ++bits;   // keys are 0..n_groups (n_groups = "dropped")
  const int passes = (bits + kRadixBits - 1) / kRadixBits;
  const unsigned *kin = nullptr, *pin = nullptr;
  unsigned *kout = w.keys_a, *pout = w.pos_a;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = pass * kRadixBits;
    DET_LAUNCH(radix_hist_kernel, (int)nblocks, kRadixThreads, 0, s, idx, kin, n, ng, shift, w.hist, nblocks, n_dev);
    DET_LAUNCH(radix_binscan_kernel, kRadixBins * 32 / kRadixThreads, kRadixThreads, 0, s, w.hist, nblocks, w.totals);
    DET_LAUNCH(radix_scatter_kernel, (int)nblocks, kRadixThreads, 0, s, idx, kin, pin, n, ng, shift, w.hist, nblocks, w.totals,
               kout, pout, n_dev);
    kin = kout;
    pin = pout;
    kout = (kout == w.keys_a) ? w.keys_b : w.keys_a;
    pout = (pout == w.pos_a) ? w.pos_b : w.pos_a;
  }
  DET_LAUNCH(group_starts_kernel, (int)((n + 1 + 255) / 256), 256, 0, s, kin, n, ng, w.starts, ngd, n_dev, w.n_long);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool vec4 = (dim % 4 == 0) && ((((uintptr_t)rows | (uintptr_t)out) & 15u) == 0);
  // rows that can be cut into 16-column slices (the double-buffered cp.async staging moves 16 B vectors)
  const int sliceable = (vec4 && dim % kSliceCols == 0 && dim > (size_t)kSliceCols) ? 1 : 0;
  DET_LAUNCH(group_long_kernel, (int)((n_groups + 255) / 256), 256, 0, s, w.starts, ng, sliceable, w.long_list, w.huge_list,
             w.n_long, ngd);
  unsigned vpr, lpr, sh;
  fgeom((unsigned)dim, vec4, 1, &vpr, &lpr, &sh);
  const unsigned gpw = 32u >> sh;
  // CTAs of the launch reserved for the long groups (default one per SM); a tuning knob until the kernel has been timed
  int n_long_ctas = env_int("DET_SEGRED_LONG_CTAS", 2 * sms);
  if (n_long_ctas < 1) n_long_ctas = 1;
  if (n_long_ctas > 4 * sms) n_long_ctas = 4 * sms;
  const auto k4 = segment_reduce_kernel<4>;
  const auto k1 = segment_reduce_kernel<1>;
  const int occ = vec4 ? occupancy_of(k4, kThreadsF) : occupancy_of(k1, kThreadsF);
  // one resident wave: the long-path CTAs (blockIdx < n_long_ctas) plus the short-group CTAs that still fit beside them
  const int long_per_sm = (n_long_ctas + sms - 1) / sms;
  const int grid = n_long_ctas + grid_for(n_groups, (int)(gpw * 2 * (kThreadsF / 32)), sms, occ > long_per_sm ? occ - long_per_sm : 1);
  if (vec4)
    DET_LAUNCH(k4, grid, kThreadsF, 0, s, rows, pin, w.starts, ng, (unsigned)dim, vpr, lpr, sh, w.long_list, w.huge_list, w.n_long, n_long_ctas, out, ngd);
  else
    DET_LAUNCH(k1, grid, kThreadsF, 0, s, rows, pin, w.starts, ng, (unsigned)dim, vpr, lpr, sh, w.long_list, w.huge_list, w.n_long, n_long_ctas, out, ngd);
  CUDA_TRY(cudaGetLastError());
  return DET_OK;
}


det_status det_segment_reduce(const float* rows, const int32_t* idx, size_t n, size_t n_groups, size_t dim, float* out,
                              void* workspace, size_t workspace_bytes, det_stream_t stream) {
  return seg_reduce_run(rows, idx, n, n_groups, dim, out, workspace, workspace_bytes, nullptr, (cudaStream_t)stream);
}

// ---- sparse apply_gradients with REPEATED ids, one call, no host synchronisation -----------------------------------
// The reference's optimizer patch handles IndexedSlices gradients with `_resource_apply_sparse_duplicate_indices`:
// unique(ids) -> unsorted_segment_sum(grads, idx, n_unique) -> find / dense rule / upsert per unique id
// (python/ops/dynamic_embedding_optimizer.py:150,184 + :161-204).  Here: det_unique -> det_segment_reduce (position
// order) -> the fused find-or-insert optimizer kernel, chained on the caller's stream with the unique count staying on
// the DEVICE (kernels read it from memory; grids and buffers are sized for the bound n), so a training step has no
// cudaStreamSynchronize and no `.item()`.  New keys start from the broadcast `init_param` row (a per-key initializer
// needs the unique ids on the host side first: use det_unique + det_segment_reduce + det_apply_*).
struct DupWs {
  long long* uniq;
  int* idx;
  long long* cnt;
  unsigned char* uws;
  size_t uws_bytes;
  float* gsum;
  unsigned char* sws;
  size_t sws_bytes;
};
static size_t dup_layout(size_t n, size_t dim, unsigned char* base, DupWs* w) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* p = base ? base + off : nullptr;
    off += align256(bytes);
    return p;
  };
  const size_t ub = unique_ws_layout(n, nullptr, nullptr), sb = seg_reduce_layout(n, n, nullptr, nullptr);
  unsigned char* uniq = take(n * 8);
  unsigned char* idx = take(n * 4);
  unsigned char* cnt = take(16);
  unsigned char* uws = take(ub);
  unsigned char* gsum = take(n * dim * 4);
  unsigned char* sws = take(sb);
  if (w) {
    w->uniq = (long long*)uniq; w->idx = (int*)idx; w->cnt = (long long*)cnt; w->uws = uws; w->uws_bytes = ub;
    w->gsum = (float*)gsum; w->sws = sws; w->sws_bytes = sb;
  }
  return off;
}

size_t det_apply_dup_workspace_bytes(size_t n, size_t dim) { return dup_layout(n ? n : 1, dim ? dim : 1, nullptr, nullptr); }

static det_status apply_dup(det_table* t, const int64_t* ids, const float* grads, size_t n, OptHyper h,
                            const float* init_param, int opt, void* workspace, size_t workspace_bytes,
                            int64_t* n_unique_dev_out, cudaStream_t s, const long long* n_items_dev = nullptr) {
  if (!t) return fail(DET_INVALID_ARGUMENT, "det_apply_dup: null table");
  if (n == 0) return DET_OK;
  if (!ids || !grads || !init_param || !workspace) return fail(DET_INVALID_ARGUMENT, "det_apply_dup: null argument");
  if (n >= 0x7fffffffull) return fail(DET_INVALID_ARGUMENT, "det_apply_dup: n too large");
  const size_t dim = (size_t)t->cfg.dim;
  if (workspace_bytes < det_apply_dup_workspace_bytes(n, dim)) return fail(DET_INVALID_ARGUMENT, "det_apply_dup: workspace too small");
  if (((uintptr_t)workspace & 255u) != 0) return fail(DET_INVALID_ARGUMENT, "det_apply_dup: workspace must be 256 B aligned");
  det::DevGuard _dg(t->cfg.device);
  DupWs w;
  dup_layout(n, dim, (unsigned char*)workspace, &w);
  // n_items_dev: the id count itself lives on the device (owner side of the sharded backward); n is its bound
  det_status st = unique_run(ids, n, (int64_t*)w.uniq, w.idx, (int64_t*)w.cnt, w.uws, w.uws_bytes, n_items_dev, s);
  if (st != DET_OK) return st;
  st = seg_reduce_run(grads, w.idx, n, n, dim, w.gsum, w.sws, w.sws_bytes, w.cnt, s, n_items_dev);
  if (st != DET_OK) return st;
  st = apply_common(t, (const int64_t*)w.uniq, w.gsum, n, h, init_param, 0, opt, s, w.cnt);
  if (st != DET_OK) return st;
  if (n_unique_dev_out) CUDA_TRY(cudaMemcpyAsync(n_unique_dev_out, w.cnt, sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  return DET_OK;
}

det_status det_apply_adagrad_dup(det_table* t, const int64_t* ids, const float* grads, size_t n, float lr, float epsilon,
                                 const float* init_param, float init_accum, void* workspace, size_t workspace_bytes,
                                 int64_t* n_unique_dev_out, det_stream_t stream) {
  OptHyper h;
  h.lr = lr; h.eps = epsilon; h.beta1 = 0.f; h.beta2 = 0.f; h.init_slot = init_accum;
  if (t) t->slot_init[1] = init_accum;
  return apply_dup(t, ids, grads, n, h, init_param, 0, workspace, workspace_bytes, n_unique_dev_out, (cudaStream_t)stream);
}

det_status det_apply_adam_dup(det_table* t, const int64_t* ids, const float* grads, size_t n, float alpha, float beta1,
                              float beta2, float epsilon, const float* init_param, void* workspace, size_t workspace_bytes,
                              int64_t* n_unique_dev_out, det_stream_t stream) {
  OptHyper h;
  h.lr = alpha; h.eps = epsilon; h.beta1 = beta1; h.beta2 = beta2; h.init_slot = 0.f;
  return apply_dup(t, ids, grads, n, h, init_param, 1, workspace, workspace_bytes, n_unique_dev_out, (cudaStream_t)stream);
}

}  // extern "C"

namespace det {
// owner side of the sharded backward (sharded.cu): ids / grads were compacted out of the inbox segments by a kernel,
// their count lives in *n_items_dev; opt 0 = Adagrad (lr, eps, init_slot), 1 = Adam (lr = alpha, eps, beta1, beta2)
det_status apply_dup_on_device_count(det_table* t, const int64_t* ids, const float* grads, size_t n_bound,
                                     const long long* n_items_dev, int opt, float lr, float eps, float beta1, float beta2,
                                     float init_slot, const float* init_param, void* workspace, size_t workspace_bytes,
                                     cudaStream_t s) {
  OptHyper h;
  h.lr = lr; h.eps = eps; h.beta1 = beta1; h.beta2 = beta2; h.init_slot = init_slot;
  if (t && opt == 0) t->slot_init[1] = init_slot;
  return apply_dup(t, ids, grads, n_bound, h, init_param, opt, workspace, workspace_bytes, nullptr, s, n_items_dev);
}
}  // namespace det

