// cuda_emu.h -- a minimal SIMT emulator for TESTING device code without a GPU (test infrastructure only; nothing in
// the product includes it).  The kernel sources are compiled unchanged by g++ with -DDET_EMU:
//   * every WARP is an OS thread and every lane a fiber (ucontext) of it; a block's warps run concurrently, blocks
//     run one after the other (so `__shared__` can simply be `static`);
//   * warp collectives (__shfl_sync, __ballot_sync, __any_sync, __syncwarp) park the lane until all live lanes of
//     the warp have arrived and exchange values through a per-warp mailbox -- divergence needs no special care
//     because each lane has its own stack and program counter; __syncthreads parks it until the block has arrived;
//   * atomics are the GCC __atomic builtins on ordinary host memory, so races between warps are REAL races
//     between OS threads: the claim / repair protocols are exercised under genuine concurrency.
// Only full-mask collectives are supported (all this code base uses).
#pragma once
#ifndef DET_EMU
#define DET_EMU 1
#endif
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <ucontext.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) __attribute__((aligned(x)))

struct emu_dim3 {
  unsigned x = 0, y = 0, z = 0;
};

struct longlong2 {
  long long x, y;
};
static inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }
struct alignas(16) int4 {
  int x, y, z, w;
};
struct alignas(8) int2 {
  int x, y;
};
struct alignas(16) float4 {
  float x, y, z, w;
};
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace emu {
// One OS thread per WARP; its 32 lanes are fibers (ucontext) scheduled round-robin by that thread.  A lane that
// reaches a warp collective parks until all live lanes of the warp have arrived; __syncthreads parks it until all
// live threads of the block have.  Warps of a block run truly concurrently (atomics, CAS races are real), blocks run
// one after the other (so `__shared__` can be `static`).
constexpr size_t kFiberStack = 256 * 1024;

struct Fiber {
  ucontext_t ctx;
  emu_dim3 tid;
  int lane = 0;
  bool done = false;
  void* stack = nullptr;
};

struct Block {
  std::atomic<unsigned> arrived{0};
  std::atomic<unsigned> gen{0};
  std::atomic<unsigned> alive{0};
};

struct Warp {
  Fiber f[32];
  int n_lanes = 0, alive = 0, cur = 0;
  ucontext_t sched;
  unsigned long long vals[32] = {};
  int arrived = 0;
  unsigned gen = 0;
  Block* block = nullptr;
  const void* fn = nullptr;
  void (*invoke)(const void*) = nullptr;
};

inline thread_local Warp* W = nullptr;
inline thread_local emu_dim3 t_blockIdx, t_blockDim, t_gridDim;

static inline Fiber* self() { return &W->f[W->cur]; }
static inline void yield_lane() { swapcontext(&W->f[W->cur].ctx, &W->sched); }

// barrier over the live lanes of the warp
static inline void warp_barrier() {
  Warp* w = W;
  const unsigned my = w->gen;
  if (++w->arrived >= w->alive) {
    w->arrived = 0;
    ++w->gen;
  } else {
    while (w->gen == my) yield_lane();
  }
}

static inline void block_barrier() {
  Block* b = W->block;
  const unsigned my = b->gen.load();
  if (b->arrived.fetch_add(1) + 1 >= b->alive.load()) {
    b->arrived.store(0);
    b->gen.fetch_add(1);
  } else {
    while (b->gen.load() == my) yield_lane();
  }
}

inline unsigned long long now_ns() {
  return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

inline void require_full(unsigned mask) {
  if (mask != 0xffffffffu) {
    fprintf(stderr, "cuda_emu: only full-mask warp collectives are supported\n");
    abort();
  }
}

static void fiber_main() {
  Warp* w = W;
  Fiber* me = &w->f[w->cur];
  w->invoke(w->fn);
  // an exited thread no longer takes part in barriers: release whoever waits for it
  me->done = true;
  w->vals[me->lane] = 0;
  --w->alive;
  if (w->alive > 0 && w->arrived >= w->alive) {
    w->arrived = 0;
    ++w->gen;
  }
  Block* b = w->block;
  const unsigned left = b->alive.fetch_sub(1) - 1;
  if (left > 0 && b->arrived.load() >= left) {
    b->arrived.store(0);
    b->gen.fetch_add(1);
  }
  swapcontext(&me->ctx, &w->sched);
}

static void run_warp(Warp* w, emu_dim3 bi, emu_dim3 bd, emu_dim3 gd) {
  W = w;
  t_blockIdx = bi;
  t_blockDim = bd;
  t_gridDim = gd;
  for (int l = 0; l < w->n_lanes; ++l) {
    Fiber& f = w->f[l];
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kFiberStack;
    f.ctx.uc_link = nullptr;
    makecontext(&f.ctx, (void (*)())fiber_main, 0);
  }
  int idle_rounds = 0;
  while (w->alive > 0) {
    bool progressed = false;
    for (int l = 0; l < w->n_lanes; ++l) {
      if (w->f[l].done) continue;
      const unsigned g0 = w->gen;
      const int a0 = w->arrived;
      w->cur = l;
      swapcontext(&w->sched, &w->f[l].ctx);
      if (w->f[l].done || w->gen != g0 || w->arrived != a0) progressed = true;
    }
    // every live lane is parked on the block barrier (or on an atomic another warp must release): let others run
    if (!progressed && ++idle_rounds > 4) {
      std::this_thread::yield();
      idle_rounds = 0;
    }
  }
}

// run `fn` as a grid of `grid` blocks of `block` threads (a last partial warp must not use warp collectives that
// name absent lanes, as on the GPU)
// `__shared__` variables are plain statics, so two kernels must never run at the same time: launches from different
// host threads ("ranks" of an emulated multi-rank group) are serialised.  The one kernel that WAITS for other ranks'
// kernels (the peer flag barrier; it uses no shared memory) is launched with launch_unlocked.
inline std::mutex& launch_mutex() {
  static std::mutex mu;
  return mu;
}

template <typename F>
void launch_unlocked(unsigned grid, unsigned block, F fn) {
  if (block == 0 || grid == 0) abort();
  const unsigned n_warps = (block + 31) / 32;
  std::vector<std::unique_ptr<Warp>> warps;
  for (unsigned w = 0; w < n_warps; ++w) {
    warps.emplace_back(new Warp());
    for (int l = 0; l < 32; ++l) warps[w]->f[l].stack = malloc(kFiberStack);
  }
  for (unsigned b = 0; b < grid; ++b) {
    Block blk;
    blk.alive.store(block);
    std::vector<std::thread> th;
    for (unsigned w = 0; w < n_warps; ++w) {
      Warp* wp = warps[w].get();
      wp->n_lanes = (int)(block - w * 32 < 32 ? block - w * 32 : 32);
      wp->alive = wp->n_lanes;
      wp->arrived = 0;
      wp->gen = 0;
      wp->block = &blk;
      wp->fn = &fn;
      wp->invoke = [](const void* p) { (*(const F*)p)(); };
      for (int l = 0; l < 32; ++l) {
        wp->f[l].done = l >= wp->n_lanes;
        wp->f[l].lane = l;
        wp->f[l].tid.x = w * 32 + (unsigned)l;
        wp->vals[l] = 0;
      }
      emu_dim3 bi, bd, gd;
      bi.x = b;
      bd.x = block;
      gd.x = grid;
      if (n_warps == 1)
        run_warp(wp, bi, bd, gd);
      else
        th.emplace_back(run_warp, wp, bi, bd, gd);
    }
    for (auto& x : th) x.join();
  }
  for (auto& w : warps)
    for (int l = 0; l < 32; ++l) free(w->f[l].stack);
}

template <typename F>
void launch(unsigned grid, unsigned block, F fn) {
  std::lock_guard<std::mutex> lk(launch_mutex());
  launch_unlocked(grid, block, fn);
}
}  // namespace emu

#define threadIdx (::emu::self()->tid)
#define blockIdx (::emu::t_blockIdx)
#define blockDim (::emu::t_blockDim)
#define gridDim (::emu::t_gridDim)

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
  emu::require_full(mask);
  emu::warp_barrier();
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
// a spinning lane must let the other lanes of its warp (and other warps / "ranks") run
static inline long long clock64() {
  emu::yield_lane();
  return (long long)emu::now_ns();
}

template <typename T>
static inline unsigned long long emu_bits(T v) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  unsigned long long b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <typename T>
static inline T emu_from(unsigned long long b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
// all lanes deposit, all lanes read, nobody overwrites before everybody has read
template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src) {
  emu::require_full(mask);
  emu::Warp& w = *emu::W;
  w.vals[emu::self()->lane] = emu_bits(v);
  emu::warp_barrier();
  const unsigned long long r = w.vals[src & 31];
  emu::warp_barrier();
  return emu_from<T>(r);
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
  emu::require_full(mask);
  emu::Warp& w = *emu::W;
  const int lane = emu::self()->lane;
  w.vals[lane] = emu_bits(v);
  emu::warp_barrier();
  const unsigned src = (unsigned)lane + delta;
  const unsigned long long r = src < 32 ? w.vals[src] : emu_bits(v);
  emu::warp_barrier();
  return emu_from<T>(r);
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
  emu::require_full(mask);
  emu::Warp& w = *emu::W;
  const int lane = emu::self()->lane;
  w.vals[lane] = emu_bits(v);
  emu::warp_barrier();
  const unsigned long long r = (unsigned)lane >= delta ? w.vals[lane - delta] : emu_bits(v);
  emu::warp_barrier();
  return emu_from<T>(r);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
  emu::require_full(mask);
  emu::Warp& w = *emu::W;
  const int lane = emu::self()->lane;
  w.vals[lane] = emu_bits(v);
  emu::warp_barrier();
  const unsigned long long r = w.vals[(lane ^ lane_mask) & 31];
  emu::warp_barrier();
  return emu_from<T>(r);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  emu::require_full(mask);
  emu::Warp& w = *emu::W;
  w.vals[emu::self()->lane] = pred ? 1ull : 0ull;
  emu::warp_barrier();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)(w.vals[i] & 1ull) << i;
  emu::warp_barrier();
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline unsigned __match_any_sync(unsigned mask, int v) {
  emu::require_full(mask);
  emu::Warp& w = *emu::W;
  w.vals[emu::self()->lane] = emu_bits(v);
  emu::warp_barrier();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i)
    if (!w.f[i].done && w.vals[i] == emu_bits(v)) r |= 1u << i;
  emu::warp_barrier();
  return r;
}

// ---- atomics / intrinsics ---------------------------------------------------------------------------------------
static inline unsigned long long atomicCAS(unsigned long long* a, unsigned long long cmp, unsigned long long val) {
  __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}
static inline unsigned long long atomicCAS_system(unsigned long long* a, unsigned long long cmp, unsigned long long val) {
  return atomicCAS(a, cmp, val);
}
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) {
  return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST);
}
static inline unsigned long long atomicAdd_system(unsigned long long* a, unsigned long long v) { return atomicAdd(a, v); }
static inline unsigned atomicAdd_system(unsigned* a, unsigned v) { return atomicAdd(a, v); }
static inline unsigned long long atomicMin(unsigned long long* a, unsigned long long v) {
  unsigned long long cur = __atomic_load_n(a, __ATOMIC_SEQ_CST);
  while (v < cur && !__atomic_compare_exchange_n(a, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return cur;
}
static inline unsigned atomicMin(unsigned* a, unsigned v) {
  unsigned cur = __atomic_load_n(a, __ATOMIC_SEQ_CST);
  while (v < cur && !__atomic_compare_exchange_n(a, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return cur;
}
static inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v) {
  unsigned long long cur = __atomic_load_n(a, __ATOMIC_SEQ_CST);
  while (v > cur && !__atomic_compare_exchange_n(a, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
  }
  return cur;
}
static inline unsigned atomicOr(unsigned* a, unsigned v) { return __atomic_fetch_or(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr_system(unsigned* a, unsigned v) { return atomicOr(a, v); }
static inline unsigned atomicExch(unsigned* a, unsigned v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch_system(unsigned* a, unsigned v) { return atomicExch(a, v); }

template <typename T>
static inline T __ldg(const T* p) {
  return *p;
}
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
  return (unsigned long long)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __float_as_uint(float f) { return emu_from<unsigned>(emu_bits(f)); }
