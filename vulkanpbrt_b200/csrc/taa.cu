// taa.cu -- k_taa: YCoCg neighbourhood-test temporal anti-aliasing on the tone-mapped image.
//
// Replaces shaders/taa.comp:48-104 and the final->history vkCmdCopyImage that follows it
// (source/renderModules/Taa.cpp:99-143).  The copy disappears: the kernel's output buffer IS the
// next frame's history (ping-pong pair owned by the Taa module), read back through the same
// "BGRA8 bytes viewed as RGBA8" reinterpretation the raw copy produces, so the
// reference's R/B-swapped history is reproduced unless fix_swizzle is set.
//
// Streaming: 1 pixel per thread, 32x16 CTAs.  The YCoCg transform of the current image is staged
// once per texel in a shared-memory tile with a 1-texel apron and the 3x3 bounds are formed
// separably from it.  Algorithmic traffic per pixel: denoised 4 + motion 4 + history 4 read,
// final 4 written = 16 B.
// All arithmetic is non-contracted IEEE so the BGRA8 output is bit-exact against the oracle.
#include "common.cuh"
#include "kernels.h"

namespace vkpbrt {

// taa.comp:32-38 on a tone-mapped texel.  Inputs are finite and non-negative (unorm8 / 255), so the
// shader's "* 1.f" factors are identities and its "g * 0.f" term is an added +0: dropping them is exact.
VK_DEVICE void ycocg(float r, float g, float b, float& y, float& co, float& cg)
{
    y = add_rn(add_rn(r, mul_rn(g, 2.f)), b);
    co = add_rn(mul_rn(r, 2.f), mul_rn(b, -2.f));
    cg = add_rn(add_rn(-r, mul_rn(g, 2.f)), -b);
}

constexpr int TAA_BX = 32, TAA_BY = 16;                     // output pixels per CTA
constexpr int TAA_TW = TAA_BX + 2, TAA_TH = TAA_BY + 2;     // YCoCg tile with a 1-texel apron

__global__ void __launch_bounds__(TAA_BX* TAA_BY) k_taa(const TaaParams p)
{
    // YCoCg of every texel of the tile is computed ONCE (the shader converts each texel 9 times).
    // Out-of-image neighbours are skipped by the shader (taa.comp:70); clamping the coordinate instead
    // re-reads a texel that is already in the same box / cross set, so min and max are unchanged.
    __shared__ float s_y[TAA_TH][TAA_TW], s_co[TAA_TH][TAA_TW], s_cg[TAA_TH][TAA_TW];
    const int W = p.W, H = p.H;
    const int x0 = blockIdx.x * TAA_BX, y0 = p.row_begin + blockIdx.y * TAA_BY;
    const int tid = threadIdx.y * TAA_BX + threadIdx.x;
    for (int i = tid; i < TAA_TW * TAA_TH; i += TAA_BX * TAA_BY) {
        const int ty = i / TAA_TW, tx = i - ty * TAA_TW;
        int sx = x0 + tx - 1, sy = y0 + ty - 1;
        sx = sx < 0 ? 0 : (sx >= W ? W - 1 : sx);
        sy = sy < 0 ? 0 : (sy >= H ? H - 1 : sy);
        const uint32_t sb = __ldg(p.denoised + (size_t)sy * W + sx);          // BGRA8: byte0 = B
        ycocg(unorm8_byte_to_f32(sb, 2), unorm8_byte_to_f32(sb, 1), unorm8_byte_to_f32(sb, 0),
              s_y[ty][tx], s_co[ty][tx], s_cg[ty][tx]);
    }
    __syncthreads();
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;
    if (gx >= W || gy >= p.row_end) return;                                   // taa.comp:50
    const size_t pix = (size_t)gy * W + gx;
    const uint32_t cur_bits = __ldg(p.denoised + pix);
    const uint32_t mv = __ldg(p.motion + pix);
    const float u = f16_bits_to_f32((uint16_t)(mv & 0xffffu)), v = f16_bits_to_f32((uint16_t)(mv >> 16));
    if (p.frame == 0 || u < 0.0f || v < 0.0f || u > 1.0f || v > 1.0f) {          // :57-60
        p.final_bgra[pix] = cur_bits | 0xff000000u;   // unorm8 -> float -> unorm8 is the identity
        return;
    }
    // :62-83 box (3x3) and cross (5 texel) bounds, separably: per row the 3-texel min / max
    const int tx = threadIdx.x + 1, ty = threadIdx.y + 1;
    float mnb[3], mxb[3], mnc[3], mxc[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float(*pl)[TAA_TW] = c == 0 ? s_y : (c == 1 ? s_co : s_cg);
        float rmin[3], rmax[3];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const float a = pl[ty - 1 + dy][tx - 1], b = pl[ty - 1 + dy][tx], d = pl[ty - 1 + dy][tx + 1];
            rmin[dy] = fminf(fminf(a, b), d);
            rmax[dy] = fmaxf(fmaxf(a, b), d);
        }
        const float up = pl[ty - 1][tx], dn = pl[ty + 1][tx];
        mnb[c] = fminf(fminf(rmin[0], rmin[1]), rmin[2]);
        mxb[c] = fmaxf(fmaxf(rmax[0], rmax[1]), rmax[2]);
        mnc[c] = fminf(fminf(rmin[1], up), dn);
        mxc[c] = fmaxf(fmaxf(rmax[1], up), dn);
    }
    // :88 bilinear history.  Bytes are B,G,R,A of the previous final; sampled as RGBA8 the .x
    // channel reads byte 0.
    const Bilin bl = bilin_setup(u, v, W, H);
    const uint32_t h00 = __ldg(p.history + (size_t)bl.y0 * W + bl.x0), h10 = __ldg(p.history + (size_t)bl.y0 * W + bl.x1);
    const uint32_t h01 = __ldg(p.history + (size_t)bl.y1 * W + bl.x0), h11 = __ldg(p.history + (size_t)bl.y1 * W + bl.x1);
    float prev[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int k = p.fix_swizzle ? (2 - c) : c;
        prev[c] = bilin_mix(bl, unorm8_byte_to_f32(h00, k), unorm8_byte_to_f32(h10, k), unorm8_byte_to_f32(h01, k), unorm8_byte_to_f32(h11, k));
    }
    // history is a bilinear mix of unorm8 values: finite and non-negative as well
    float pyc[3];
    ycocg(prev[0], prev[1], prev[2], pyc[0], pyc[1], pyc[2]);
    bool inside = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float mn = mul_rn(add_rn(mnb[c], mnc[c]), .5f), mx = mul_rn(add_rn(mxb[c], mxc[c]), .5f);
        if (!(pyc[c] >= mn) || !(pyc[c] <= mx)) inside = false;
    }
    if (!inside) {
        p.final_bgra[pix] = cur_bits | 0xff000000u;
        return;
    }
    const float cur[3] = {unorm8_byte_to_f32(cur_bits, 2), unorm8_byte_to_f32(cur_bits, 1), unorm8_byte_to_f32(cur_bits, 0)};
    float res[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) res[c] = add_rn(mul_rn(.4f, cur[c]), mul_rn((1 - .4f), prev[c]));
    p.final_bgra[pix] = (uint32_t)f32_to_unorm8(res[2]) | ((uint32_t)f32_to_unorm8(res[1]) << 8) |
                        ((uint32_t)f32_to_unorm8(res[0]) << 16) | 0xff000000u;
}

cudaError_t launch_taa(const TaaParams& p, cudaStream_t stream)
{
    const int rows = p.row_end - p.row_begin;
    if (rows <= 0) return cudaSuccess;
    dim3 block(TAA_BX, TAA_BY, 1), grid((p.W + TAA_BX - 1) / TAA_BX, (rows + TAA_BY - 1) / TAA_BY, 1);
    VKPBRT_LAUNCH(k_taa, grid, block, 0, stream, p);
    return cudaGetLastError();
}

}  // namespace vkpbrt
