// taa.cu -- k_taa: YCoCg neighbourhood-test temporal anti-aliasing on the tone-mapped image.
//
// Replaces shaders/taa.comp:48-104 and the final->history vkCmdCopyImage that follows it
// (source/renderModules/Taa.cpp:99-143).  The copy disappears: the kernel's output buffer IS the
// next frame's history (ping-pong pair owned by the Taa module), read back through the same
// "BGRA8 bytes viewed as RGBA8" reinterpretation the raw copy produces, so the
// reference's R/B-swapped history is reproduced unless fix_swizzle is set.
//
// Streaming, HBM-bound: 1 pixel per thread, 32x8 CTAs; the 3x3 neighbourhood of the current image
// is served by L1 (each texel is reused by 9 threads of the same CTA).  Algorithmic traffic per
// pixel: denoised 4 + motion 4 + history 4 read, final 4 written = 16 B.
// All arithmetic is non-contracted IEEE so the BGRA8 output is bit-exact against the oracle.
#include "common.cuh"
#include "kernels.h"

namespace vkpbrt {

// taa.comp:32-38
VK_DEVICE void ycocg(float r, float g, float b, float* o)
{
    o[0] = add_rn(add_rn(mul_rn(r, 1.f), mul_rn(g, 2.f)), mul_rn(b, 1.f));
    o[1] = add_rn(add_rn(mul_rn(r, 2.f), mul_rn(g, 0.f)), mul_rn(b, -2.f));
    o[2] = add_rn(add_rn(mul_rn(r, -1.f), mul_rn(g, 2.f)), mul_rn(b, -1.f));
}

__global__ void __launch_bounds__(256) k_taa(const TaaParams p)
{
    const int gx = blockIdx.x * 32 + threadIdx.x;
    const int gy = p.row_begin + blockIdx.y * 8 + threadIdx.y;
    if (gx >= p.W || gy >= p.row_end) return;                                   // taa.comp:50
    const int W = p.W, H = p.H;
    const size_t pix = (size_t)gy * W + gx;
    const uint32_t cur_bits = __ldg(p.denoised + pix);                          // BGRA8: byte0 = B
    const float cur[3] = {unorm8_to_f32((cur_bits >> 16) & 0xffu), unorm8_to_f32((cur_bits >> 8) & 0xffu),
                          unorm8_to_f32(cur_bits & 0xffu)};
    const uint32_t mv = __ldg(p.motion + pix);
    const float u = f16_bits_to_f32((uint16_t)(mv & 0xffffu)), v = f16_bits_to_f32((uint16_t)(mv >> 16));
    if (p.frame == 0 || u < 0.0f || v < 0.0f || u > 1.0f || v > 1.0f) {          // :57-60
        p.final_bgra[pix] = cur_bits | 0xff000000u;   // unorm8 -> float -> unorm8 is the identity
        return;
    }
    // vec3(1/0) folds to INT_MAX in glslang; the centre texel always replaces it
    float mnb[3], mnc[3], mxb[3], mxc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mnb[c] = mnc[c] = 2147483647.0f;
        mxb[c] = mxc[c] = -2147483647.0f;
    }
#pragma unroll
    for (int y = -1; y <= 1; ++y)
#pragma unroll
        for (int x = -1; x <= 1; ++x) {
            const int sx = gx + x, sy = gy + y;
            if (sx >= 0 && sy >= 0 && sx < W && sy < H) {                       // :70
                const uint32_t sb = (x == 0 && y == 0) ? cur_bits : __ldg(p.denoised + (size_t)sy * W + sx);
                float yc[3];
                ycocg(unorm8_to_f32((sb >> 16) & 0xffu), unorm8_to_f32((sb >> 8) & 0xffu), unorm8_to_f32(sb & 0xffu), yc);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (x == 0 || y == 0) {
                        mnc[c] = gl_min(mnc[c], yc[c]);
                        mxc[c] = gl_max(mxc[c], yc[c]);
                    }
                    mnb[c] = gl_min(mnb[c], yc[c]);
                    mxb[c] = gl_max(mxb[c], yc[c]);
                }
            }
        }
    // :88 bilinear history.  Bytes are B,G,R,A of the previous final; sampled as RGBA8 the .x
    // channel reads byte 0.
    const Bilin bl = bilin_setup(u, v, W, H);
    const uint32_t h00 = __ldg(p.history + (size_t)bl.y0 * W + bl.x0), h10 = __ldg(p.history + (size_t)bl.y0 * W + bl.x1);
    const uint32_t h01 = __ldg(p.history + (size_t)bl.y1 * W + bl.x0), h11 = __ldg(p.history + (size_t)bl.y1 * W + bl.x1);
    float prev[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int sh = 8 * (p.fix_swizzle ? (2 - c) : c);
        prev[c] = bilin_mix(bl, unorm8_to_f32((h00 >> sh) & 0xffu), unorm8_to_f32((h10 >> sh) & 0xffu),
                            unorm8_to_f32((h01 >> sh) & 0xffu), unorm8_to_f32((h11 >> sh) & 0xffu));
    }
    float pyc[3];
    ycocg(prev[0], prev[1], prev[2], pyc);
    bool inside = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float mn = mul_rn(add_rn(mnb[c], mnc[c]), .5f), mx = mul_rn(add_rn(mxb[c], mxc[c]), .5f);
        if (!(pyc[c] >= mn) || !(pyc[c] <= mx)) inside = false;
    }
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) res[c] = inside ? add_rn(mul_rn(.4f, cur[c]), mul_rn((1 - .4f), prev[c])) : cur[c];
    p.final_bgra[pix] = (uint32_t)f32_to_unorm8(res[2]) | ((uint32_t)f32_to_unorm8(res[1]) << 8) |
                        ((uint32_t)f32_to_unorm8(res[0]) << 16) | 0xff000000u;
}

cudaError_t launch_taa(const TaaParams& p, cudaStream_t stream)
{
    const int rows = p.row_end - p.row_begin;
    if (rows <= 0) return cudaSuccess;
    dim3 block(32, 8, 1), grid((p.W + 31) / 32, (rows + 7) / 8, 1);
    VKPBRT_LAUNCH(k_taa, grid, block, 0, stream, p);
    return cudaGetLastError();
}

}  // namespace vkpbrt
