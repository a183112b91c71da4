// host.h -- host-side table object and shared declarations for the .cu translation units.
#pragma once
#ifndef DET_EMU  // the emulator build gets its own definitions of these (tests/emu/cuda_emu.h, cuda_runtime_emu.h)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <stdlib.h>

#include <mutex>
#include <shared_mutex>
#include <string>
#include <type_traits>

#include "../../include/detable.h"
#include "common.cuh"

// Kernel launch.  The test suite's SIMT emulator (tests/emu/, -DDET_EMU) compiles these translation units with g++
// and runs the launch as OS threads; nvcc sees the ordinary <<<>>> launch.
#ifdef DET_EMU
#define DET_LAUNCH(kernel, grid, block, smem, stream, ...) \
  ::emu::launch((unsigned)(grid), (unsigned)(block), [&] { kernel(__VA_ARGS__); })
#define DET_LAUNCH_SPIN(kernel, grid, block, smem, stream, ...) \
  ::emu::launch_unlocked((unsigned)(grid), (unsigned)(block), [&] { kernel(__VA_ARGS__); })
#define DET_DYN_SHARED(name) static __attribute__((aligned(16))) unsigned char name[96 * 1024]
#else
#define DET_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// a kernel that waits for kernels of OTHER processes (the peer flag barrier): same launch on the GPU
#define DET_LAUNCH_SPIN(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DET_DYN_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#endif

namespace det {

extern thread_local std::string g_last_error;
det_status fail(det_status code, const std::string& msg);

#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      cudaGetLastError();                                                                           \
      return ::det::fail(_e == cudaErrorMemoryAllocation ? DET_OUT_OF_MEMORY : DET_CUDA_ERROR,      \
                         std::string("CUDA error at " __FILE__ ":") + std::to_string(__LINE__) +    \
                             ": " + cudaGetErrorString(_e));                                        \
    }                                                                                               \
  } while (0)

// resident CTAs per SM of a kernel (cached per kernel): grids are sized to exactly one resident wave
template <typename K>
inline int occupancy_of(K kernel, int threads) {
  static thread_local const void* last_k = nullptr;
  static thread_local int last_v = 0;
  if (last_k == (const void*)kernel) return last_v;
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0) != cudaSuccess || nb < 1) {
    cudaGetLastError();
    nb = 2;
  }
  last_k = (const void*)kernel;
  last_v = nb;
  return nb;
}

inline int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Every entry point runs on its table's device and puts the caller's current device back on return: the
// caller's framework (torch, TF) must never find its current CUDA context changed by a table call.
struct DevGuard {
  int prev = -1;
  bool changed = false;
  explicit DevGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) {
      cudaGetLastError();
      prev = -1;
    }
    if (prev != dev) {
      cudaSetDevice(dev);
      changed = true;
    }
  }
  ~DevGuard() {
    if (changed && prev >= 0) cudaSetDevice(prev);
  }
  DevGuard(const DevGuard&) = delete;
  DevGuard& operator=(const DevGuard&) = delete;
};

struct HostPipe;    // chunked H2D -> kernel -> D2H pipeline state (host_api.cu)
struct EvictState;  // score plane + eviction bookkeeping of a table with an eviction strategy (evict.cu)

size_t dtype_size(int dt);
RowGeom make_geom(unsigned row_bytes, int vec);
int pick_vec(size_t row_bytes, const void* a, const void* b, const void* c);
int grid_for(size_t n_items, int items_per_block, int sm_count, int blocks_per_sm);

}  // namespace det

struct det_table {
  // host-side bookkeeping of mutating / scratch-using entry points (the reference takes a mutex for mutators and a
  // shared lock for readers, hkv_hashtable_op_gpu.cu.cc:201-364; det_find needs no host lock)
  std::mutex mu;
  // Readers that take no `mu` (det_find, det_find_scores: no bookkeeping, CUDA-graph capturable) hold this SHARED while
  // they copy `view` and enqueue their kernel; rehash_to holds it EXCLUSIVELY around "device-wide sync, free the old
  // planes, swap view".  A find launched from another host thread therefore either completed before the old planes
  // are freed or already sees the new ones (the reference: tf_shared_lock for readers, hkv_hashtable_op_gpu.cu.cc:201).
  std::shared_mutex view_mu;
  det_config cfg;
  size_t row_bytes = 0;
  float max_lf = 0.75f;
  int sm_count = 148;
  det::TableView view{};
  void* raw[1 + det::kMaxPlanes];
  det::DevState* h_state = nullptr;  // pinned host mirror
  uint64_t used_ub = 0;              // host upper bound of non-EMPTY slots (no sync needed)
  unsigned long long* h_used_snap = nullptr;  // pinned: async snapshot of DevState::used
  cudaEvent_t snap_ev = nullptr;
  bool snap_inflight = false;
  uint64_t n_since_snap = 0;         // keys of mutating calls issued after the snapshot in flight
  uint32_t rehash_count = 0;
  float slot_init[det::kMaxPlanes];  // value given to slot-plane rows of keys created by insert/accum
  det::HostPipe* pipe = nullptr;   // host-buffer pipeline of det_find_host
  det::HostPipe* pipe2 = nullptr;  // ... of det_insert_host (separate streams: the two directions overlap)
  void* scratch = nullptr;           // per-table device scratch reused by det_lookup_sparse / det_export
  size_t scratch_bytes = 0;
  bool external = false;             // planes live in a caller-provided region (not owned, fixed capacity)
  unsigned long long* peer_bar = nullptr;  // arrival flags of the NVLink peer barrier (sharded.cu)
  det::EvictState* ev = nullptr;     // non-null: the table has an eviction strategy (evict.cu)
  uint64_t last_used_snap = 0;       // last exact value of DevState::used the host has seen (snapshot or sync read)
  void* sparse_flag = nullptr;       // det_lookup_sparse: where the one-id-per-row flag word lives (scratch may move)
  unsigned sparse_epoch = 0;         // ... and the call counter it is compared with
};

namespace det {
struct RegionLayout {
  uint64_t nb;
  size_t off_state, off_bar, off_keys, off_plane[kMaxPlanes], bytes;
};
void region_layout(const det_config& cfg, RegionLayout* L);
struct SlotInit;
det_status ensure_room(det_table* t, const long long* keys_or_null, size_t n, cudaStream_t s);
void note_mutation(det_table* t, size_t n, cudaStream_t s);
det_status insert_impl(det_table* t, const int64_t* keys, const void* values, size_t n, cudaStream_t s,
                       bool check_room);
det_status table_clear_async(det_table* t, cudaStream_t s);
// stream-ordered users only (one stream per table at a time); grows with a sync, never shrinks
det_status table_scratch(det_table* t, size_t bytes, void** out);
void host_pipe_free(det_table* t);
det_status rehash_to(det_table* t, uint64_t new_nb, cudaStream_t s);
// fused.cu: unique -> position-order gradient sum -> fused optimizer step with the id count on the device (sharded.cu)
det_status apply_dup_on_device_count(det_table* t, const int64_t* ids, const float* grads, size_t n_bound,
                                     const long long* n_items_dev, int opt, float lr, float eps, float beta1, float beta2,
                                     float init_slot, const float* init_param, void* workspace, size_t workspace_bytes,
                                     cudaStream_t s);
// evict.cu (all called with t->mu held, except evict_insert which takes it)
det_status evict_attach(det_table* t, int strategy);
void evict_free(det_table* t);
bool evict_at_max(const det_table* t);
det_status evict_room(det_table* t, const long long* keys_or_null, size_t n, cudaStream_t s);
det_status evict_on_rehash(det_table* t, const TableView& ov, const TableView& nv, cudaStream_t s);
void evict_on_clear(det_table* t, cudaStream_t s);
det_status evict_before_remove(det_table* t, const long long* keys, size_t n, cudaStream_t s);
det_status evict_touch(det_table* t, const long long* keys, const unsigned long long* scores, size_t n, cudaStream_t s);
det_status evict_insert(det_table* t, const int64_t* keys, const void* values, const uint64_t* scores, size_t n,
                        cudaStream_t s);
void evict_stats(const det_table* t, uint32_t* events, uint64_t* evicted);
unsigned long long* evict_scores(const det_table* t);  // score plane [capacity + 2] (t->ev != nullptr)
ScoreRule evict_rule(const det_table* t);              // strategy + current epoch
}  // namespace det
