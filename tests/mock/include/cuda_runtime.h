// TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A synchronous, host-memory stand-in for the handful of CUDA runtime calls the HOST side of
// libmodelxdigest.so makes (modelx_b200/csrc/mxd_api.cu, mxd_lockstep.cu, mxd_hasher.cu).  tests/mock_build.py
// compiles those same sources against this header plus tests/mock/mock_kernels.cpp (kernel launchers that hash on
// the CPU with the oracle) into tests/mock/_build/libmodelxdigest_mock.so, so the host logic -- the digest
// service's rounds, windows and chains, per-call cancellation, fd limits, sinks, the C++ client mirror, the CLI --
// can be exercised in the CPU-only container (`-m "not gpu"`), also under ThreadSanitizer.  The product never
// loads this library: modelx_b200 loads modelx_b200/libmodelxdigest.so, and the parity tests proper (`-m gpu`)
// run the real CUDA build.  "Device" memory is malloc'ed host memory tracked in a registry so
// cudaPointerGetAttributes can tell the kinds apart; streams and events are no-ops because every call completes
// before it returns.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNoDevice = 100, cudaErrorNotReady = 600 };
typedef struct mock_stream* cudaStream_t;
typedef struct mock_event* cudaEvent_t;
typedef void* cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocPortable = 1, cudaHostRegisterPortable = 1, cudaHostRegisterReadOnly = 8 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };

namespace mockcuda {
struct Registry {
    std::mutex mu;
    std::map<uintptr_t, std::pair<size_t, int>> blocks;   // base -> (size, kind: 1 host-pinned, 2 device) ; device ordinal in upper bits
    int current = 0;
};
Registry& registry();
int device_count();
void* alloc(size_t n, int kind, int device);
void release(void* p);
int lookup(const void* p, int* device);    // 0 unregistered, 1 pinned host, 2 device
int& current_device();
}  // namespace mockcuda

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorNoDevice ? "no CUDA-capable device is detected (mock)" : "mock CUDA error"); }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = mockcuda::device_count(); return *n > 0 ? cudaSuccess : cudaErrorNoDevice; }
inline cudaError_t cudaGetDevice(int* d) { *d = mockcuda::current_device(); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= mockcuda::device_count()) return cudaErrorInvalidValue; mockcuda::current_device() = d; return cudaSuccess; }
inline cudaError_t cudaDeviceGetPCIBusId(char*, int, int) { return cudaErrorInvalidValue; }
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }

inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = reinterpret_cast<cudaStream_t>(new char); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete reinterpret_cast<char*>(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = reinterpret_cast<cudaEvent_t>(new char); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete reinterpret_cast<char*>(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = static_cast<T*>(mockcuda::alloc(n, 2, mockcuda::current_device())); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <typename T> inline cudaError_t cudaMallocAsync(T** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
inline cudaError_t cudaFree(void* p) { mockcuda::release(p); return cudaSuccess; }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { mockcuda::release(p); return cudaSuccess; }
template <typename T> inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { *p = static_cast<T*>(mockcuda::alloc(n, 1, 0)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFreeHost(void* p) { mockcuda::release(p); return cudaSuccess; }
inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    int dev = 0;
    const int kind = mockcuda::lookup(p, &dev);
    a->type = kind == 2 ? cudaMemoryTypeDevice : (kind == 1 ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered);
    a->device = dev; a->devicePointer = const_cast<void*>(p); a->hostPointer = const_cast<void*>(p);
    return cudaSuccess;
}
