// hostsim.h -- TEST-ONLY SIMT emulator: runs the CUDA kernel sources of vulkanpbrt_b200/csrc on the
// CPU so the kernels' logic can be debugged against the oracle in the GPU-less dev container.
//
// NOT a product path and NOT a fallback: nothing in the vulkanpbrt_b200 package can load the library
// built from this header; only tests/conftest.py does (marker "hostsim"), and its version string
// says so.  GPU parity (tests -m gpu) is always measured on the real libvkpbrt_b200.so.
//
// How: every CUDA thread of a block is a fiber (hand-rolled x86-64 context switch).  A fiber runs
// until it reaches __syncthreads() or a warp shuffle; the scheduler resolves warp exchanges as soon
// as every live lane of a warp has arrived and releases a block barrier when every live thread
// has.  Blocks are distributed over OpenMP threads.  __shared__ becomes static thread_local.
#pragma once
#ifndef VKPBRT_HOSTSIM
#error "hostsim.h is only for the test emulator build"
#endif

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <time.h>

#include <functional>
#include <vector>

// ---- CUDA vocabulary --------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __constant__ static const

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(8) uint2 { uint32_t x, y; };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct alignas(8) float2 { float x, y; };
inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct uchar4 { uint8_t x, y, z, w; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801 };
typedef struct hostsim_stream_s* cudaStream_t;
enum { cudaStreamNonBlocking = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
struct cudaUUID_t { char bytes[16]; };
struct cudaDeviceProp { char name[64]; cudaUUID_t uuid; int major, minor; };

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "hostsim error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { strcpy(p->name, "HOSTSIM (CPU test emulator)"); memcpy(p->uuid.bytes, "HOSTSIM-DEVICE-0", 16); p->major = 10; p->minor = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
// hostsim.cpp.  HOSTSIM_GUARD=end|start places every allocation against an inaccessible page (its end, 16-byte granular,
// or its start): a kernel that reads or writes outside a plane faults instead of touching a neighbouring allocation.
cudaError_t cudaMalloc(void** p, size_t n);
cudaError_t cudaFree(void* p);
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
// launches are synchronous: events and stream order are no-ops
typedef struct hostsim_event_s* cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
// "peer" memory: the ranks of a multi-rank test are THREADS of one process, so a handle is the pointer itself
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { memset(h, 0, sizeof(*h)); memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
// atomics / fences / timer used by the halo kernels (ranks = OS threads: real concurrency)
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicMin(uint32_t* p, uint32_t v)
{
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
    return old;
}
inline uint32_t atomicExch(uint32_t* p, uint32_t v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
    return old;
}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __nanosleep(unsigned) { sched_yield(); }
// external memory / semaphores (tests/vkmock plays the Vulkan side): an exported "VkDeviceMemory" is a memfd that is
// mapped shared, an exported semaphore a memfd holding one 64-bit payload.  As with CUDA, a successful import takes
// ownership of the file descriptor.  Launches are synchronous, so a wait spins (bounded) until the payload arrives.
struct hostsim_external_memory_s;
struct hostsim_external_semaphore_s;
typedef hostsim_external_memory_s* cudaExternalMemory_t;
typedef hostsim_external_semaphore_s* cudaExternalSemaphore_t;
enum { cudaExternalMemoryHandleTypeOpaqueFd = 1, cudaExternalSemaphoreHandleTypeOpaqueFd = 1, cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd = 9 };
enum { cudaExternalMemoryDedicated = 1 };
struct cudaExternalMemoryHandleDesc { int type; struct { int fd; } handle; unsigned long long size; unsigned flags; };
struct cudaExternalMemoryBufferDesc { unsigned long long offset, size; unsigned flags; };
struct cudaExternalSemaphoreHandleDesc { int type; struct { int fd; } handle; unsigned flags; };
struct cudaExternalSemaphoreWaitParams { struct { struct { unsigned long long value; } fence; } params; unsigned flags; };
struct cudaExternalSemaphoreSignalParams { struct { struct { unsigned long long value; } fence; } params; unsigned flags; };
cudaError_t cudaImportExternalMemory(cudaExternalMemory_t*, const cudaExternalMemoryHandleDesc*);
cudaError_t cudaExternalMemoryGetMappedBuffer(void**, cudaExternalMemory_t, const cudaExternalMemoryBufferDesc*);
cudaError_t cudaDestroyExternalMemory(cudaExternalMemory_t);
cudaError_t cudaImportExternalSemaphore(cudaExternalSemaphore_t*, const cudaExternalSemaphoreHandleDesc*);
cudaError_t cudaWaitExternalSemaphoresAsync(cudaExternalSemaphore_t*, const cudaExternalSemaphoreWaitParams*, unsigned, cudaStream_t);
cudaError_t cudaSignalExternalSemaphoresAsync(cudaExternalSemaphore_t*, const cudaExternalSemaphoreSignalParams*, unsigned, cudaStream_t);
cudaError_t cudaDestroyExternalSemaphore(cudaExternalSemaphore_t);

// ---- fibers -----------------------------------------------------------------------------------------
extern "C" void hostsim_switch(void** save_sp, void* load_sp);

namespace hostsim {

enum State { RUN = 0, WARP_WAIT = 1, BLOCK_WAIT = 2, DONE = 3 };

struct Fiber {
    void* sp;
    int state;
    uint3 tid;
    int lane;
    uint32_t xin, xout;   // warp exchange mailbox
    int xsrc;
};

struct Block {
    dim3 grid, block;
    uint3 bid;
    Fiber* fibers;
    int n;
    Fiber* cur;
    void* sched_sp;
    const std::function<void()>* body;
};

Block*& blk();                       // per OS thread
unsigned char* dyn_smem();           // per OS thread, 256 KB (dynamic shared memory of the running block)
void launch(dim3 grid, dim3 block, const std::function<void()>& body);
void yield(int new_state);

inline uint32_t warp_exchange(uint32_t v, int src_lane)
{
    Fiber* f = blk()->cur;
    f->xin = v;
    f->xsrc = src_lane;
    yield(WARP_WAIT);
    return f->xout;
}

}  // namespace hostsim

#define threadIdx (hostsim::blk()->cur->tid)
#define blockIdx (hostsim::blk()->bid)
#define blockDim (hostsim::blk()->block)
#define gridDim (hostsim::blk()->grid)

#define VKPBRT_LAUNCH(kernel, grid, block, smem, stream, ...) \
    hostsim::launch((grid), (block), [&]() { kernel(__VA_ARGS__); })

inline void __syncthreads() { hostsim::yield(hostsim::BLOCK_WAIT); }

inline float __shfl_xor_sync(unsigned, float v, int mask)
{
    uint32_t u;
    memcpy(&u, &v, 4);
    u = hostsim::warp_exchange(u, hostsim::blk()->cur->lane ^ mask);
    memcpy(&v, &u, 4);
    return v;
}
inline float __shfl_sync(unsigned, float v, int src)
{
    uint32_t u;
    memcpy(&u, &v, 4);
    u = hostsim::warp_exchange(u, src & 31);
    memcpy(&v, &u, 4);
    return v;
}

inline int __all_sync(unsigned, int pred)
{
    uint32_t v = pred ? 1u : 0u;
    for (int off = 16; off >= 1; off >>= 1) v &= hostsim::warp_exchange(v, hostsim::blk()->cur->lane ^ off);
    return (int)v;
}

// ---- intrinsics ---------------------------------------------------------------------------------------
// (the emulator build uses -ffp-contract=off, so plain operators are the _rn forms)
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __frcp_rn(float a) { return 1.0f / a; }
template <typename T>
inline T __ldg(const T* p) { return *p; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

struct __half { uint16_t bits; };
inline __half __float2half_rn(float f)
{
    uint32_t x;
    memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u, ax = x & 0x7fffffffu, h;
    if (ax > 0x7f800000u) h = 0x7fffu;
    else if (ax >= 0x477ff000u) h = 0x7c00u;
    else if (ax < 0x38800000u) {
        if (ax <= 0x33000000u) h = 0;
        else {
            uint32_t e = ax >> 23, m = (ax & 0x7fffffu) | 0x800000u, s = 126u - e;
            h = m >> s;
            uint32_t rem = m & ((1u << s) - 1u), half = 1u << (s - 1u);
            if (rem > half || (rem == half && (h & 1u))) h++;
        }
    } else {
        uint32_t e = (ax >> 23) - 112u, m = ax & 0x7fffffu;
        h = (e << 10) | (m >> 13);
        uint32_t rem = m & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
    }
    return __half{(uint16_t)(sign | h)};
}
inline float __half2float(__half hh)
{
    uint16_t h = hh.bits;
    uint32_t sign = ((uint32_t)h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu, u;
    if (e == 0) {
        float v = (float)m * 5.9604644775390625e-08f;
        return sign ? -v : v;
    }
    if (e == 31) u = sign | 0x7f800000u | (m << 13);
    else u = sign | ((e + 112u) << 23) | (m << 13);
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline uint16_t __half_as_ushort(__half h) { return h.bits; }
inline __half __ushort_as_half(uint16_t b) { return __half{b}; }
