// rt.h — the few runtime calls the host orchestration needs (device memory, streams, events, launches).
// Product build: CUDA runtime.  Test-only build (TSGPU_SIMT, tests/simt/): the fiber emulator, so the same
// orchestration code can be exercised on the GPU-less build box.  libtsgpu.so is never built that way.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string>

#ifdef TSGPU_SIMT
#include "simt.h"
#include <cstdlib>
#include <cstring>
namespace rt {
typedef void* stream_t;
typedef int event_t;
inline int device_count() { return 1; }
inline const char* set_device(int) { return nullptr; }
inline bool poison() { static const bool on = getenv("TSGPU_SIMT_POISON") && atoi(getenv("TSGPU_SIMT_POISON")) != 0; return on; }
inline const char* malloc_device(void** p, size_t n) {      // cudaMalloc does not zero memory: TSGPU_SIMT_POISON=1 fills it with 0xCD
    *p = calloc(1, n + 64);
    if (*p && poison()) memset(*p, 0xCD, n + 64);
    return *p ? nullptr : "calloc failed";
}
inline void free_device(void* p) { free(p); }
inline const char* malloc_host(void** p, size_t n) { *p = calloc(1, n + 64); return *p ? nullptr : "calloc failed"; }
inline void free_host(void* p) { free(p); }
inline const char* h2d(void* d, const void* s, size_t n, stream_t) { memcpy(d, s, n); return nullptr; }
inline const char* d2h(void* d, const void* s, size_t n, stream_t) { memcpy(d, s, n); return nullptr; }
inline const char* d2d(void* d, const void* s, size_t n, stream_t) { memmove(d, s, n); return nullptr; }
inline const char* memset_async(void* d, int v, size_t n, stream_t) { memset(d, v, n); return nullptr; }
inline const char* stream_create(stream_t* s) { *s = nullptr; return nullptr; }
inline void stream_destroy(stream_t) {}
inline const char* stream_sync(stream_t) { return nullptr; }
inline const char* event_create(event_t* e) { *e = 0; return nullptr; }
inline void event_destroy(event_t) {}
inline const char* event_record(event_t, stream_t) { return nullptr; }
inline const char* event_sync(event_t) { return nullptr; }
inline const char* stream_wait_event(stream_t, event_t) { return nullptr; }
inline const char* last_error() { return nullptr; }
typedef int tevent_t;
inline void tevent_create(tevent_t*) {}
inline void tevent_destroy(tevent_t) {}
inline void tevent_record(tevent_t, stream_t) {}
inline float tevent_ms(tevent_t, tevent_t) { return 0.f; }
inline const char* device_sync() { return nullptr; }
template <class K> inline const char* allow_smem(K, size_t) { return nullptr; }
}  // namespace rt
// the emulator is single-threaded by construction: concurrent host threads take turns launching
#include <mutex>
namespace rt { inline std::mutex& simt_launch_mutex() { static std::mutex m; return m; } }
#define TS_LAUNCH(kern, grid, block, smem, stream, ...) \
    do { std::lock_guard<std::mutex> simt_lock_(rt::simt_launch_mutex()); simt::launch(grid, block, smem, [&] { kern(__VA_ARGS__); }); } while (0)
#else
#include <cuda_runtime.h>
namespace rt {
typedef cudaStream_t stream_t;
typedef cudaEvent_t event_t;
inline const char* err(cudaError_t e) { return e == cudaSuccess ? nullptr : cudaGetErrorString(e); }
inline int device_count() { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
inline const char* set_device(int d) { return err(cudaSetDevice(d)); }
inline const char* malloc_device(void** p, size_t n) { return err(cudaMalloc(p, n)); }
inline void free_device(void* p) { if (p) cudaFree(p); }
inline const char* malloc_host(void** p, size_t n) { return err(cudaHostAlloc(p, n, cudaHostAllocPortable)); }
inline void free_host(void* p) { if (p) cudaFreeHost(p); }
inline const char* h2d(void* d, const void* s, size_t n, stream_t st) { return err(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st)); }
inline const char* d2h(void* d, const void* s, size_t n, stream_t st) { return err(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st)); }
inline const char* d2d(void* d, const void* s, size_t n, stream_t st) { return err(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st)); }
inline const char* memset_async(void* d, int v, size_t n, stream_t st) { return err(cudaMemsetAsync(d, v, n, st)); }
inline const char* stream_create(stream_t* s) { return err(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking)); }
inline void stream_destroy(stream_t s) { if (s) cudaStreamDestroy(s); }
inline const char* stream_sync(stream_t s) { return err(cudaStreamSynchronize(s)); }
inline const char* event_create(event_t* e) { return err(cudaEventCreateWithFlags(e, cudaEventDisableTiming)); }
inline void event_destroy(event_t e) { if (e) cudaEventDestroy(e); }
inline const char* event_record(event_t e, stream_t s) { return err(cudaEventRecord(e, s)); }
inline const char* event_sync(event_t e) { return err(cudaEventSynchronize(e)); }
inline const char* stream_wait_event(stream_t s, event_t e) { return err(cudaStreamWaitEvent(s, e, 0)); }
inline const char* last_error() { return err(cudaGetLastError()); }
typedef cudaEvent_t tevent_t;
inline void tevent_create(tevent_t* e) { cudaEventCreate(e); }
inline void tevent_destroy(tevent_t e) { if (e) cudaEventDestroy(e); }
inline void tevent_record(tevent_t e, stream_t s) { cudaEventRecord(e, s); }
inline float tevent_ms(tevent_t a, tevent_t b) { float ms = 0.f; cudaEventElapsedTime(&ms, a, b); return ms; }
inline const char* device_sync() { return err(cudaDeviceSynchronize()); }
template <class K> inline const char* allow_smem(K kern, size_t bytes) {
    return err(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}
}  // namespace rt
#define TS_LAUNCH(kern, grid, block, smem, stream, ...) \
    do { kern<<<grid, block, smem, stream>>>(__VA_ARGS__); } while (0)
#endif
