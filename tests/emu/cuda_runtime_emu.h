// cuda_runtime_emu.h -- the handful of CUDA runtime calls the engine's HOST code makes, on host memory, for the
// SIMT-emulation build of the test suite (see cuda_emu.h).  "Device" memory is malloc'ed host memory, streams are
// synchronous (a launch returns when the kernel has finished), there is one device with one SM.
#pragma once
#include <cstdlib>
#include <cstring>

typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaError_t { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600,
                   cudaErrorPeerAccessAlreadyEnabled = 704, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaLimit { cudaLimitMaxL2FetchGranularity = 5 };
enum { cudaEventDisableTiming = 2, cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
struct cudaDeviceProp {
  int multiProcessorCount = 1;
  char name[64] = "simt-emulator";
};

// cudaMalloc memory is NOT zeroed on a GPU: poison it here, so that code relying on zero-initialised device memory
// fails in the emulator instead of working by the luck of fresh zero pages
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  *p = malloc(n ? n : 1);
  if (*p) memset(*p, 0xCB, n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) {
  free(p);
  return cudaSuccess;
}
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
// unified memory (value plane with an HBM budget): one host allocation, placement advice is a no-op
enum { cudaMemAttachGlobal = 1, cudaCpuDeviceId = -1 };
enum cudaMemoryAdvise { cudaMemAdviseSetPreferredLocation = 3, cudaMemAdviseSetAccessedBy = 5 };
static inline cudaError_t cudaMallocManaged(void** p, size_t n, unsigned = cudaMemAttachGlobal) { return cudaMalloc(p, n); }
static inline cudaError_t cudaMemAdvise(const void*, size_t, cudaMemoryAdvise, int) { return cudaSuccess; }
static inline cudaError_t cudaMemPrefetchAsync(const void*, size_t, int, cudaStream_t = 0) { return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) {
  memset(p, v, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) {
  return cudaMemcpy(d, s, n, k);
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetDevice(int* d) {
  *d = 0;
  return cudaSuccess;
}
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) {
  *n = 1;
  return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  *p = cudaDeviceProp();
  return cudaSuccess;
}
static inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
  *v = 1;
  return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) {
  *e = malloc(1);
  return cudaSuccess;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { return cudaEventCreateWithFlags(e, 0); }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
  *ms = 0.f;
  return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  free(e);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = malloc(1);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
  free(s);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
  cudaMemoryType type;
};
// every buffer is host memory here; report it as pinned so that the asynchronous host entry points can be exercised
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) {
  a->type = cudaMemoryTypeHost;
  return cudaSuccess;
}
// CUDA IPC: all "ranks" of an emulated group live in one process, the handle simply carries the pointer
struct cudaIpcMemHandle_t {
  char reserved[64];
};
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  memset(h, 0, sizeof(*h));
  memcpy(h->reserved, &p, sizeof(p));
  return cudaSuccess;
}
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  memcpy(p, h.reserved, sizeof(*p));
  return cudaSuccess;
}
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) {
  *can = 1;
  return cudaSuccess;
}

template <typename K>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* nb, K, int, size_t) {
  *nb = 2;
  return cudaSuccess;
}
template <typename K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) {
  return cudaSuccess;
}

// ---- 16-bit float types of the accum kernels: storage + one rounded add, enough for the emulated tests ----------
struct __half {
  unsigned short x;
};
struct __nv_bfloat16 {
  unsigned short x;
};
static inline float emu_half_to_float(__half h) {
  const unsigned s = (h.x >> 15) & 1u, e = (h.x >> 10) & 0x1fu, m = h.x & 0x3ffu;
  unsigned bits;
  if (e == 0) {
    if (m == 0) {
      bits = s << 31;
    } else {
      int ee = -1;
      unsigned mm = m;
      do {
        ++ee;
        mm <<= 1;
      } while (!(mm & 0x400u));
      bits = (s << 31) | ((unsigned)(127 - 15 - ee) << 23) | ((mm & 0x3ffu) << 13);
    }
  } else if (e == 31) {
    bits = (s << 31) | 0x7f800000u | (m << 13);
  } else {
    bits = (s << 31) | ((e + 127 - 15) << 23) | (m << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
static inline __half emu_float_to_half(float f) {  // round to nearest even
  unsigned bits;
  memcpy(&bits, &f, 4);
  const unsigned s = (bits >> 16) & 0x8000u;
  int e = (int)((bits >> 23) & 0xffu) - 127 + 15;
  unsigned m = bits & 0x7fffffu;
  __half h;
  if (((bits >> 23) & 0xffu) == 0xffu) {
    h.x = (unsigned short)(s | 0x7c00u | (m ? 0x200u : 0));
  } else if (e >= 31) {
    h.x = (unsigned short)(s | 0x7c00u);
  } else if (e <= 0) {
    if (e < -10) {
      h.x = (unsigned short)s;
    } else {
      m |= 0x800000u;
      const int shift = 14 - e;
      unsigned r = m >> shift;
      const unsigned rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
      if (rem > half || (rem == half && (r & 1u))) ++r;
      h.x = (unsigned short)(s | r);
    }
  } else {
    unsigned r = ((unsigned)e << 10) | (m >> 13);
    const unsigned rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;
    h.x = (unsigned short)(s | r);
  }
  return h;
}
static inline __half __hadd(__half a, __half b) { return emu_float_to_half(emu_half_to_float(a) + emu_half_to_float(b)); }
static inline float emu_bf16_to_float(__nv_bfloat16 h) {
  const unsigned bits = (unsigned)h.x << 16;
  float f;
  memcpy(&f, &bits, 4);
  return f;
}
static inline __nv_bfloat16 __hadd(__nv_bfloat16 a, __nv_bfloat16 b) {
  const float f = emu_bf16_to_float(a) + emu_bf16_to_float(b);
  unsigned bits;
  memcpy(&bits, &f, 4);
  const unsigned lsb = (bits >> 16) & 1u;
  bits += 0x7fffu + lsb;
  __nv_bfloat16 r;
  r.x = (unsigned short)(bits >> 16);
  return r;
}
