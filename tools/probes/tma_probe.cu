// minimal TMA probe: 2-D box load of a float plane through cp.async.bulk.tensor, descriptor inside a by-value struct
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
struct alignas(64) TmaDesc { unsigned char bytes[128]; };
struct Params { int W, H; float* out; int pad; TmaDesc d; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ Params p, const __grid_constant__ CUtensorMap m2)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float* tile = reinterpret_cast<float*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4096);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(4096) : "memory");
        const void* desc = MODE == 0 ? (const void*)&p.d : (const void*)&m2;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(tile)), "l"(desc), "r"(32 * (int)blockIdx.x + p.pad), "r"(32 * (int)blockIdx.y + p.pad), "r"(smem_u32(bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra W2;\nbra W1;\nW2:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        const int r = i / 32, c = i % 32;
        const int gy = blockIdx.y * 32 + r + p.pad, gx = blockIdx.x * 32 + c + p.pad;
        if (gx < p.W && gy < p.H) p.out[(size_t)gy * p.W + gx] = tile[i];
    }
}
int main()
{
    const int W = 256, H = 128;
    std::vector<float> h(W * H);
    for (int i = 0; i < W * H; ++i) h[i] = (float)i;
    float *src, *dst;
    cudaMalloc(&src, W * H * 4); cudaMalloc(&dst, W * H * 4);
    cudaMemcpy(src, h.data(), W * H * 4, cudaMemcpyHostToDevice);
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W * 4};
    const cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    CUresult r = ((encode_fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d, sizeof(CUtensorMap)=%zu alignof=%zu\n", (int)r, sizeof(CUtensorMap), alignof(CUtensorMap));
    Params p{}; p.W = W; p.H = H; p.out = dst; memcpy(p.d.bytes, &m, 128);
    for (int mode = 0; mode < 4; ++mode) {
        p.pad = mode >= 2 ? 5 : 0;          // modes 2, 3: box origin at an arbitrary (not 16-byte aligned) element
        cudaMemcpy(dst, src, W * H * 4, cudaMemcpyDeviceToDevice);
        if (mode % 2 == 0) k<0><<<dim3(W / 32 - (mode >= 2), H / 32 - (mode >= 2)), 128, 8192>>>(p, m); else k<1><<<dim3(W / 32 - (mode >= 2), H / 32 - (mode >= 2)), 128, 8192>>>(p, m);
        e = cudaDeviceSynchronize();
        std::vector<float> o(W * H);
        cudaMemcpy(o.data(), dst, W * H * 4, cudaMemcpyDeviceToHost);
        int bad = 0; for (int i = 0; i < W * H; ++i) bad += o[i] != h[i];
        printf("mode %d (%s): %s, mismatches %d\n", mode, (mode & 1) ? "CUtensorMap param" : "descriptor inside a struct param", cudaGetErrorString(e), bad);
        if (e != cudaSuccess) break;
    }
    return 0;
}
