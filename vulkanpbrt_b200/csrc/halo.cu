// Band-sharded multi-GPU runs: halo rows pushed straight into the neighbour's
// HBM over NVLink peer mappings, with flag words for ordering.  The reference renders on one
// device and has no counterpart; what is exchanged is derived from its shaders' read footprints
// (accumulator.comp:75-98 history taps, bmfrPost.comp:103-118 / taa.comp:44-60 neighbourhoods),
// see vulkanpbrt_b200/multigpu.py BandPlan.
//
//   k_halo_push   [announce "ready" to the senders] -> [gate: spin until the local "ready" words reach gate_value] -> copy every block of rows of the table
//                 with 16-byte peer stores -> the last CTA publishes `value` to the receivers'
//                 "done" flags (release at system scope).  One launch per exchange point.
//   k_halo_wait   one warp; lane i spins (acquire at system scope) until flag i >= value.
//
// Flags only grow (value = frame + 1).  Every spin is bounded: on a timeout the error word is set
// and the kernel returns, so a lost peer surfaces as an error at the next flush instead of a hung GPU.
#include "common.cuh"
#include "kernels.h"

namespace vkpbrt {

namespace {

#ifndef VKPBRT_HOSTSIM
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#else   // test emulator: ranks are OS threads of one process
inline uint32_t ld_acquire_sys(const uint32_t* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_sys(uint32_t* p, uint32_t v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned long long global_timer_ns()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
}
#endif

// returns the nanoseconds spent spinning
__device__ unsigned long long spin_until(const uint32_t* flag, uint32_t value, unsigned long long timeout_ns, uint32_t* error)
{
    if ((int32_t)(ld_acquire_sys(flag) - value) >= 0) return 0ull;
    const unsigned long long t0 = global_timer_ns();
    while ((int32_t)(ld_acquire_sys(flag) - value) < 0) {
        if (global_timer_ns() - t0 > timeout_ns) {
            atomicExch(error, 1u);
            break;
        }
        __nanosleep(40);
    }
    return global_timer_ns() - t0;
}

template <typename V>
__device__ __forceinline__ void copy_block(const HaloCopy& hc, uint32_t rows, uint32_t row_bytes, int part, int parts)
{
    const uint32_t vpr = row_bytes / (uint32_t)sizeof(V);
    const uint32_t total = rows * vpr;
    const uint32_t stride = (uint32_t)parts * blockDim.x;
    uint32_t i = (uint32_t)part * blockDim.x + threadIdx.x;
    // four independent loads in flight per thread before the (posted) peer stores
    for (; i + 3 * stride < total; i += 4 * stride) {
        V v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t j = i + k * stride, r = j / vpr, c = j - r * vpr;
            v[k] = *reinterpret_cast<const V*>(hc.src + (size_t)r * hc.src_pitch + (size_t)c * sizeof(V));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t j = i + k * stride, r = j / vpr, c = j - r * vpr;
            *reinterpret_cast<V*>(hc.dst + (size_t)r * hc.dst_pitch + (size_t)c * sizeof(V)) = v[k];
        }
    }
    for (; i < total; i += stride) {
        const uint32_t r = i / vpr, c = i - r * vpr;
        *reinterpret_cast<V*>(hc.dst + (size_t)r * hc.dst_pitch + (size_t)c * sizeof(V)) =
            *reinterpret_cast<const V*>(hc.src + (size_t)r * hc.src_pitch + (size_t)c * sizeof(V));
    }
}

}  // namespace

// grid = (parts per block of rows, blocks of rows in the table)
__global__ void __launch_bounds__(256) k_halo_push(const HaloPushParams p)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < p.n_announce) st_release_sys(p.announce_flags[threadIdx.x], p.value);
    if (p.n_ready > 0) {
        if ((int)threadIdx.x < p.n_ready) {
            const unsigned long long ns = spin_until(p.ready_flags[threadIdx.x], p.gate_value, p.timeout_ns, p.error);
            if (ns && blockIdx.x == 0 && blockIdx.y == 0) atomicMax(p.gate_ns + 1, ns), atomicAdd(p.gate_ns, ns);
        }
        __syncthreads();
    }
    if ((int)blockIdx.y < p.n_copies) {
        const HaloCopy hc = p.copies[blockIdx.y];
        uint32_t rows = hc.rows, row_bytes = hc.row_bytes;
        if (hc.src_pitch == row_bytes && hc.dst_pitch == row_bytes) {      // contiguous on both sides: one long row
            row_bytes *= rows;
            rows = 1;
        }
        const uintptr_t align = (uintptr_t)hc.src | (uintptr_t)hc.dst | row_bytes |
                                (rows > 1 ? (uintptr_t)(hc.src_pitch | hc.dst_pitch) : 0);
        if ((align & 15) == 0) copy_block<uint4>(hc, rows, row_bytes, blockIdx.x, gridDim.x);
        else if ((align & 3) == 0) copy_block<uint32_t>(hc, rows, row_bytes, blockIdx.x, gridDim.x);
        else copy_block<uint8_t>(hc, rows, row_bytes, blockIdx.x, gridDim.x);
    }
    // completion: every CTA fences its stores to system scope, the last one to arrive publishes the flags
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        if (atomicAdd(p.counter, 1u) == total - 1) {
            *p.counter = 0;
            __threadfence_system();
            for (int i = 0; i < p.n_done; ++i) st_release_sys(p.done_flags[i], p.value);
        }
    }
}

__global__ void __launch_bounds__(32) k_halo_wait(const HaloWaitParams p)
{
    if ((int)threadIdx.x < p.n) {
        const unsigned long long ns = spin_until(p.flags[threadIdx.x], p.value, p.timeout_ns, p.error);
        if (ns) atomicMax(p.wait_ns + 1, ns), atomicAdd(p.wait_ns, ns);
    }
}

cudaError_t launch_halo_push(const HaloPushParams& p, int parts, cudaStream_t stream)
{
    dim3 grid((unsigned)parts, (unsigned)(p.n_copies > 0 ? p.n_copies : 1));
    VKPBRT_LAUNCH(k_halo_push, grid, dim3(256, 1, 1), 0, stream, p);
    return cudaGetLastError();
}

cudaError_t launch_halo_wait(const HaloWaitParams& p, cudaStream_t stream)
{
    VKPBRT_LAUNCH(k_halo_wait, dim3(1, 1, 1), dim3(32, 1, 1), 0, stream, p);
    return cudaGetLastError();
}

}  // namespace vkpbrt
