// debug.cu -- device-side self checks behind the C ABI's vkpbrt_debug_* entry points (used by tests/ only).
#include "common.cuh"
#include "kernels.h"

namespace vkpbrt {

// every non-negative binary32 (0 .. +inf): the threshold search against the vk_pow form of the tone-map quantiser
__global__ void __launch_bounds__(256) k_tonemap_sweep(unsigned long long* bad, uint32_t* first_bad)
{
    __shared__ uint32_t thr[256];
    thr[threadIdx.x] = c_tonemap_thr[threadIdx.x];
    __syncthreads();
    const uint32_t end = 0x7f800000u;       // +inf, inclusive
    unsigned long long mine = 0;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= end; b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((uint32_t)b);
        if (tonemap_code(x, thr) != tonemap_code_reference(x)) {
            ++mine;
            atomicMin(first_bad, (uint32_t)b);
        }
    }
    if (mine) atomicAdd(bad, mine);
}

cudaError_t launch_tonemap_sweep(unsigned long long* bad, uint32_t* first_bad, cudaStream_t stream)
{
    VKPBRT_LAUNCH(k_tonemap_sweep, dim3(148 * 16, 1, 1), dim3(256, 1, 1), 0, stream, bad, first_bad);
    return cudaGetLastError();
}

}  // namespace vkpbrt
