// cuda_emu.h -- a minimal SIMT emulator for TESTING device code without a GPU (test infrastructure only; nothing in
// the product includes it).  The kernel sources are compiled unchanged by g++ with -DDET_EMU:
//   * every CUDA thread is an OS thread; a block's threads run concurrently, blocks run one after the other
//     (so `__shared__` can simply be `static`);
//   * warp collectives (__shfl_sync, __ballot_sync, __any_sync, __syncwarp) are barriers over the 32 lane threads of
//     a warp that exchange values through a per-warp mailbox -- divergence needs no special care because each lane
//     really is its own thread; __syncthreads is a barrier over the block;
//   * atomics are the GCC __atomic builtins on ordinary host memory, so races between "CUDA threads" are REAL races
//     between OS threads: the claim / repair protocols are exercised under genuine concurrency.
// Only full-mask collectives are supported (all this code base uses).
#pragma once
#ifndef DET_EMU
#define DET_EMU 1
#endif
#include <atomic>
#include <barrier>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) alignas(x)

struct emu_dim3 {
  unsigned x = 0, y = 0, z = 0;
};
inline thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

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

namespace emu {
struct Warp {
  std::barrier<> bar{32};
  unsigned long long vals[32] = {};
};
inline thread_local Warp* warp = nullptr;
inline thread_local int lane = 0;
inline thread_local std::barrier<>* block_bar = nullptr;

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

// run `fn` as a grid of `grid` blocks of `block` threads (block % 32 == 0)
template <typename F>
void launch(unsigned grid, unsigned block, F fn) {
  if (block % 32 != 0) abort();
  for (unsigned b = 0; b < grid; ++b) {
    std::barrier<> bbar((std::ptrdiff_t)block);
    std::vector<std::unique_ptr<Warp>> warps;
    for (unsigned w = 0; w < block / 32; ++w) warps.emplace_back(new Warp());
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t) {
      th.emplace_back([&, t, b]() {
        threadIdx.x = t;
        blockIdx.x = b;
        blockDim.x = block;
        gridDim.x = grid;
        warp = warps[t / 32].get();
        lane = (int)(t & 31);
        block_bar = &bbar;
        fn();
        // an exited thread no longer takes part in barriers
        warp->vals[lane] = 0;
        warp->bar.arrive_and_drop();
        bbar.arrive_and_drop();
      });
    }
    for (auto& x : th) x.join();
  }
}
}  // namespace emu

static inline void __syncthreads() { emu::block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) {
  emu::require_full(mask);
  emu::warp->bar.arrive_and_wait();
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

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
  emu::Warp& w = *emu::warp;
  w.vals[emu::lane] = emu_bits(v);
  w.bar.arrive_and_wait();
  const unsigned long long r = w.vals[src & 31];
  w.bar.arrive_and_wait();
  return emu_from<T>(r);
}
template <typename T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
  emu::require_full(mask);
  emu::Warp& w = *emu::warp;
  w.vals[emu::lane] = emu_bits(v);
  w.bar.arrive_and_wait();
  const unsigned src = (unsigned)emu::lane + delta;
  const unsigned long long r = src < 32 ? w.vals[src] : emu_bits(v);
  w.bar.arrive_and_wait();
  return emu_from<T>(r);
}
template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
  emu::require_full(mask);
  emu::Warp& w = *emu::warp;
  w.vals[emu::lane] = emu_bits(v);
  w.bar.arrive_and_wait();
  const unsigned long long r = (unsigned)emu::lane >= delta ? w.vals[emu::lane - delta] : emu_bits(v);
  w.bar.arrive_and_wait();
  return emu_from<T>(r);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  emu::require_full(mask);
  emu::Warp& w = *emu::warp;
  w.vals[emu::lane] = pred ? 1ull : 0ull;
  w.bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)(w.vals[i] & 1ull) << i;
  w.bar.arrive_and_wait();
  return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }

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
static inline unsigned long long atomicMin(unsigned long long* a, unsigned long long v) {
  unsigned long long cur = __atomic_load_n(a, __ATOMIC_SEQ_CST);
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
