// cuda_rt_emu.h — the handful of CUDA runtime calls csrc/sweep.cu and csrc/gather.cu make, on host memory, so that
// the whole host runtime (staging, drain, tick, read, run_ticks, ...) can be compiled with
// cuda_emu.h into tests/emu/libamsweep_emu.so and driven through the real C-ABI without a GPU.
// One "device"; every stream operation completes before the call returns (a legal, maximally
// serialised schedule); events are wall-clock timestamps.  Test infrastructure only.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2 };
struct EmuStream { int unused; };
struct EmuEvent { std::chrono::steady_clock::time_point t; };
typedef EmuStream* cudaStream_t;
typedef EmuEvent* cudaEvent_t;

static inline const char* cudaGetErrorString(cudaError_t e) {
  return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "emulated CUDA error");
}
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// sixteen "devices": the exchange tests run one rank per device id, as OS threads of one process
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 16; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d >= 0 && d < 16 ? cudaSuccess : cudaErrorInvalidValue; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }  // 4 "SMs"

// CUDA IPC between "processes" that are threads here: the handle carries the pointer itself
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  std::memset(h, 0, sizeof *h);
  std::memcpy(h->reserved, &p, sizeof p);
  return cudaSuccess;
}
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  std::memcpy(p, h.reserved, sizeof *p);
  return *p ? cudaSuccess : cudaErrorInvalidValue;
}
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }

static inline cudaError_t cudaMalloc(void** p, size_t bytes) {
  *p = std::aligned_alloc(256, (bytes + 255) / 256 * 256 + 256);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }

static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { return cudaMemset(d, v, n); }

static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new EmuStream{0}; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent{std::chrono::steady_clock::now()}; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  if (*ms <= 0.f) *ms = 1e-6f;
  return cudaSuccess;
}
