// tests/emu/mock/cuda_runtime.h -- TEST TOOL.  A stand-in for the CUDA runtime so that the product's C-ABI sources (csrc/*.cu) can be
// compiled by g++ and driven on host threads together with the emulated kernels (tests/emu/tm_cpu_emu.h): device memory is host memory,
// streams are synchronous, a kernel launch is emu_launch().  It exists to dry-run the `-m gpu` tests of code paths that have not yet
// seen a GPU (tests/emu/mock/README.md); nothing under trophic_maelstorm_b200/ can reach it.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>

struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} operator unsigned() const { return x; } };
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef struct mock_stream_ *cudaStream_t;
typedef struct mock_event_ { double t; } *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamDefault = 0, cudaStreamNonBlocking = 1 };
enum { cudaEventDefault = 0, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaDeviceProp {
    char name[256];
    int multiProcessorCount, major, minor;
    size_t sharedMemPerBlockOptin, sharedMemPerMultiprocessor, totalGlobalMem;
    int maxThreadsPerMultiProcessor;
};
struct cudaIpcMemHandle_t { char reserved[64]; };

inline int &mock_device() { static thread_local int d = 0; return d; }
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "mock CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { const char *e = getenv("TM_MOCK_DEVICES"); *n = e ? atoi(e) : 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { mock_device() = d; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int *d) { *d = mock_device(); return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
    memset(p, 0, sizeof *p);
    strcpy(p->name, "mock B200");
    p->multiProcessorCount = 148; p->major = 10; p->minor = 0; p->sharedMemPerBlockOptin = 232448; p->sharedMemPerMultiprocessor = 233472;
    p->totalGlobalMem = (size_t)8 << 30; p->maxThreadsPerMultiProcessor = 2048;
    return cudaSuccess;
}
inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = (size_t)6 << 30; *t = (size_t)8 << 30; return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n)
{
    // poison: device memory is not zeroed by cudaMalloc
    void *q = malloc(n ? n : 1);
    if (!q) return cudaErrorMemoryAllocation;
    memset(q, 0xA5, n);
    *p = (T *)q;
    return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr)
{
    for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { return cudaStreamCreate(s); }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { return cudaStreamCreate(s); }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)calloc(1, sizeof(mock_event_)); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr)
{
    e->t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
template <class K> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, K, int, size_t) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *) { memset(h, 0, sizeof *h); return cudaSuccess; }

// a launch: the kernel's own source on host threads (tm_cpu_emu.h is already included through the kernels' headers)
template <class K, class... A> inline void mock_launch(unsigned grid, unsigned block, size_t smem, K kernel, const A &...args)
{
    auto call = [&](const int &) { kernel(args...); };
    emu_launch(call, 0, grid, block, smem);
}
