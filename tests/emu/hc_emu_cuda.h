// hc_emu_cuda.h -- the handful of CUDA runtime calls hc_api.cu uses, mapped onto host memory.
// TEST HARNESS ONLY (see hc_emu.h); compiled only with -DHC_EMU.
#pragma once
#ifndef HC_EMU
#error "test build only"
#endif
#include <chrono>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
typedef void *cudaStream_t;
struct hc_emu_event { std::chrono::steady_clock::time_point t; };
typedef hc_emu_event *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };

static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { free(p); return 0; }
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
enum { cudaEventBlockingSync = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new hc_emu_event(); return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new hc_emu_event(); return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return 0;
}
