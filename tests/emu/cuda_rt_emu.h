// TEST INFRASTRUCTURE ONLY -- the subset of the CUDA runtime API that gpis_capi.cu calls, for the CPU
// emulation build (g++ -DGPIS_EMULATE): "device" memory is host memory (poisoned on allocation, as
// cudaMalloc's is undefined), copies are memcpy, streams are synchronous, events are wall-clock stamps,
// stream capture fails (the engine falls back to plain launches), IPC is unsupported, and there is one
// "device" of compute capability 10.0 with GPIS_EMU_SMS multiprocessors (default 3).
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>

#include "cuda_emu.h"

typedef struct emu_event { std::chrono::steady_clock::time_point t; }* cudaEvent_t;
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
struct cudaDeviceProp { int major, minor, multiProcessorCount; };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; };
struct cudaIpcMemHandle_t { char reserved[64]; };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaStreamCaptureModeThreadLocal = 1, cudaIpcMemLazyEnablePeerAccess = 1 };

inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime: unsupported"; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  const char* s = getenv("GPIS_EMU_SMS");
  p->major = 10;
  p->minor = 0;
  p->multiProcessorCount = s ? atoi(s) : 3;
  return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  void* q = malloc(bytes ? bytes : 16);
  if (!q) return cudaErrorMemoryAllocation;
  memset(q, 0xCD, bytes ? bytes : 16);                 // undefined contents: make reads of them visible
  *p = static_cast<T*>(q);
  return cudaSuccess;
}
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t i = 0; i < h; ++i) memmove(static_cast<char*>(d) + i * dp, static_cast<const char*>(s) + i * sp, w);
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorNotSupported; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = nullptr; return cudaErrorNotSupported; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorNotSupported; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event{std::chrono::steady_clock::now()}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}
inline cudaError_t cudaFuncSetAttribute(const void*, int, int) { return cudaSuccess; }
enum { cudaDevAttrClockRate = 13, cudaDevAttrCooperativeLaunch = 95 };
inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) { *v = attr == cudaDevAttrClockRate ? 1000000 : 0; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, const void*, int, size_t) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = cudaMemoryTypeDevice; return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
