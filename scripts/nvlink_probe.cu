// nvlink_probe.cu -- measurement tool (not part of the product): what does each way of WRITING to a peer GPU over NVLink
// achieve on this box?  Decides how the owner-side exchange (csrc/sharded.cu, det_peer_xchg_*) should move its rows.
// One process, two GPUs with peer access; buffers are plain cudaMalloc memory of the destination GPU.
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -o recommenders_addons_b200/lib/libdetprobe.so scripts/nvlink_probe.cu
//   run:    python scripts/nvlink_probe.py            (under gpurun --gpus 2)
// modes (row = 256 B):
//   0 cudaMemcpyPeerAsync                                 (copy-engine reference)
//   1 st.global.v4 stream, contiguous                    (ideal SM stores)
//   2 st.global.v4, 16 lanes per row, rows at RANDOM remote positions, data from registers   (serve_find's store pattern)
//   3 local RANDOM row read  -> remote CONTIGUOUS rows   (route_rows: gather locally, pack remotely)
//   4 local contiguous read  -> remote RANDOM rows       (serve_find with a cheap local side)
//   5 TMA bulk store shared -> remote, contiguous chunks of `chunk` bytes (staged by a TMA bulk load of local memory)
//   6 TMA bulk store of single rows at RANDOM remote positions (staged like 5)
//   7 remote RANDOM row READ -> local contiguous write   (the pull pattern of det_peer_find)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); return -1; } } while (0)

constexpr int kRow = 256;

__device__ __forceinline__ size_t perm_row(size_t r, size_t n_rows) { return (r * 0x9E3779B97F4A7C15ull + 12345ull) & (n_rows - 1); }

__global__ void k_stream(int4* __restrict__ dst, size_t n16) {
  const int4 v = make_int4(threadIdx.x, blockIdx.x, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// mode 2 / 3 / 4 / 7: 16 lanes per row, 4 rows in flight per lane
template <int MODE>
__global__ void k_rows(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, size_t n_rows) {
  const size_t lane16 = threadIdx.x & 15;
  const size_t grp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const size_t ngrp = ((size_t)gridDim.x * blockDim.x) >> 4;
  for (size_t r0 = grp * 4; r0 < n_rows; r0 += ngrp * 4) {
    int4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t r = r0 + u;
      if (r >= n_rows) continue;
      if (MODE == 2) v[u] = make_int4((int)r, 1, 2, 3);
      else {
        const size_t sr = (MODE == 3 || MODE == 7) ? perm_row(r, n_rows) : r;
        asm volatile("ld.global.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + sr * kRow + lane16 * 16));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t r = r0 + u;
      if (r >= n_rows) continue;
      const size_t dr = (MODE == 2 || MODE == 4) ? perm_row(r, n_rows) : r;
      asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(dst + dr * kRow + lane16 * 16), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w) : "memory");
    }
  }
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// mode 5 / 6: persistent CTAs; per iteration: TMA bulk load of `chunk` local bytes into a stage, then TMA bulk store(s) to
// the peer.  2 stages: the load of chunk i+1 overlaps the store of chunk i.  RANDOM = each 256 B row of the chunk goes to
// a random remote row (one bulk store per row), else one bulk store of the whole chunk.
template <bool RANDOM>
__global__ void k_bulk(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, size_t n_rows, unsigned chunk) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar[2];
  const unsigned rows_per_chunk = chunk / kRow;
  const size_t n_chunks = n_rows / rows_per_chunk;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[i])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;   // one elected thread drives the copies (the data never touches registers)
  unsigned it = 0;
  size_t c = blockIdx.x;
  auto issue_load = [&](size_t cc, int stage) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[stage])), "r"(chunk) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (size_t)stage * chunk)),
                 "l"(src + cc * chunk), "r"(chunk), "r"(smem_u32(&bar[stage])) : "memory");
  };
  if (c < n_chunks) issue_load(c, 0);
  for (; c < n_chunks; c += gridDim.x, ++it) {
    const int stage = it & 1;
    const size_t nxt = c + gridDim.x;
    if (nxt < n_chunks) {
      // the stage being refilled was read by the bulk stores of iteration it-1: wait until they have READ shared memory
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      issue_load(nxt, stage ^ 1);
    }
    unsigned ok = 0;
    const unsigned parity = (it >> 1) & 1u;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar[stage])), "r"(parity) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (RANDOM) {
      for (unsigned r = 0; r < rows_per_chunk; ++r) {
        const size_t dr = perm_row(c * rows_per_chunk + r, n_rows);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + dr * kRow), "r"(smem_u32(smem + (size_t)stage * chunk + (size_t)r * kRow)), "r"(kRow) : "memory");
      }
    } else {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c * chunk), "r"(smem_u32(smem + (size_t)stage * chunk)), "r"(chunk) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Side {
  int dev, peer;
  unsigned char *local = nullptr, *remote = nullptr;   // local: on `dev`; remote: on `peer`
  cudaStream_t st = nullptr;
  cudaEvent_t a = nullptr, b = nullptr;
};

static int launch(const Side& s, int mode, size_t bytes, unsigned chunk, int ctas_per_sm) {
  cudaSetDevice(s.dev);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s.dev);
  const size_t n_rows = bytes / kRow;
  const int grid = sms * ctas_per_sm;
  switch (mode) {
    case 0: CK(cudaMemcpyPeerAsync(s.remote, s.peer, s.local, s.dev, bytes, s.st)); break;
    case 1: k_stream<<<grid, 256, 0, s.st>>>((int4*)s.remote, bytes / 16); break;
    case 2: k_rows<2><<<grid, 256, 0, s.st>>>(s.local, s.remote, n_rows); break;
    case 3: k_rows<3><<<grid, 256, 0, s.st>>>(s.local, s.remote, n_rows); break;
    case 4: k_rows<4><<<grid, 256, 0, s.st>>>(s.local, s.remote, n_rows); break;
    case 7: k_rows<7><<<grid, 256, 0, s.st>>>(s.remote, s.local, n_rows); break;
    case 5:
    case 6: {
      const size_t smem = 2 * (size_t)chunk;
      if (mode == 5) {
        CK(cudaFuncSetAttribute(k_bulk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_bulk<false><<<grid, 32, smem, s.st>>>(s.local, s.remote, n_rows, chunk);
      } else {
        CK(cudaFuncSetAttribute(k_bulk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_bulk<true><<<grid, 32, smem, s.st>>>(s.local, s.remote, n_rows, chunk);
      }
    } break;
    default: return -1;
  }
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

// gbs_out[0] = dev0 -> dev1 rate (GB/s of payload), gbs_out[1] = dev1 -> dev0 rate when bidir (else 0)
int probe_run(int mode, size_t bytes, unsigned chunk, int ctas_per_sm, int bidir, int reps, double* gbs_out) {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) return -2;
  if (bytes & (bytes - 1)) return -3;   // power of two (perm_row)
  Side s[2];
  for (int d = 0; d < 2; ++d) {
    s[d].dev = d;
    s[d].peer = 1 - d;
    CK(cudaSetDevice(d));
    cudaError_t pe = cudaDeviceEnablePeerAccess(1 - d, 0);
    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) { fprintf(stderr, "no peer access %d->%d\n", d, 1 - d); return -4; }
    cudaGetLastError();
    CK(cudaMalloc(&s[d].local, bytes));
    CK(cudaMemset(s[d].local, 1, bytes));
    CK(cudaStreamCreate(&s[d].st));
    CK(cudaEventCreate(&s[d].a));
    CK(cudaEventCreate(&s[d].b));
  }
  for (int d = 0; d < 2; ++d) {   // the buffer the OTHER side writes into
    CK(cudaSetDevice(1 - d));
    CK(cudaMalloc(&s[d].remote, bytes));
    CK(cudaMemset(s[d].remote, 0, bytes));
  }
  for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
  const int sides = bidir ? 2 : 1;
  for (int w = 0; w < 2; ++w)
    for (int d = 0; d < sides; ++d)
      if (launch(s[d], mode, bytes, chunk, ctas_per_sm)) return -5;
  for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaDeviceSynchronize()); }
  for (int d = 0; d < sides; ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(s[d].a, s[d].st)); }
  for (int r = 0; r < reps; ++r)
    for (int d = 0; d < sides; ++d)
      if (launch(s[d], mode, bytes, chunk, ctas_per_sm)) return -5;
  for (int d = 0; d < sides; ++d) { CK(cudaSetDevice(d)); CK(cudaEventRecord(s[d].b, s[d].st)); }
  gbs_out[0] = gbs_out[1] = 0;
  for (int d = 0; d < sides; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaEventSynchronize(s[d].b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, s[d].a, s[d].b));
    gbs_out[d] = (double)bytes * reps / (ms * 1e-3) / 1e9;
  }
  for (int d = 0; d < 2; ++d) {
    cudaSetDevice(d);
    cudaFree(s[d].local);
    cudaStreamDestroy(s[d].st);
    cudaEventDestroy(s[d].a);
    cudaEventDestroy(s[d].b);
    cudaSetDevice(1 - d);
    cudaFree(s[d].remote);
  }
  return 0;
}
}
