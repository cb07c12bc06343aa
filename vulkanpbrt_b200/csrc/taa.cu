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
#include <atomic>

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

// ------------------------------------------------------------------------------------------------
// Work decomposition (round 2).  The first version (one pixel per thread, a YCoCg tile in shared memory, 9 + 5 shared
// loads and 20 min / max per channel and pixel) ran at 371 instructions per pixel with the ALU pipe 63 % busy: bound by
// instruction issue at 18 % of the HBM roofline.  Now:
//   * a thread owns a PAIR of adjacent columns and walks down a strip of TAA_ROWS rows with a three-row window in
//     registers, so every texel is converted to YCoCg exactly once and nothing goes through shared memory;
//   * the horizontal neighbours come from the adjacent lanes (two shuffles per channel and row), the three-texel row
//     minima / maxima are formed once per texel and reused by the three output rows that need them;
//   * a warp spans 64 columns and produces the inner 60 (the first and last lane only feed their neighbours), which
//     divides 1920 / 3840 / 7680 exactly;
//   * all arithmetic on the two pixels is issued as packed fp32 pairs (common.cuh), the history gather is
//     branch-free (a pixel that is not reprojected samples texel (0,0) and keeps the current colour by a select).
// Same operations in the same order per pixel as the oracle: the BGRA8 output is bit-exact.
#ifndef TAA_ROWS_PER_WARP
#define TAA_ROWS_PER_WARP 16
#endif
constexpr int TAA_ROWS = TAA_ROWS_PER_WARP;      // output rows per warp
constexpr int TAA_WARPS = 4;      // warps per CTA (side by side along x)
constexpr int TAA_COLS = 60;      // output columns per warp

VK_DEVICE uint32_t byte_as_float_bits(uint32_t texel, int k)      // 2^23 + byte k, as binary32 bits (one PRMT)
{
#ifndef VKPBRT_HOSTSIM
    return __byte_perm(texel, 0x4b000000u, 0x7650u + (uint32_t)k);
#else
    return 0x4b000000u | ((texel >> (8 * k)) & 0xffu);
#endif
}
// unorm8 -> float of byte k of two texels (common.cuh unorm8_byte_to_f32 on both lanes)
VK_DEVICE f2 unorm8_pair(const Pk& k, uint32_t ta, uint32_t tb, int byte)
{
    const f2 cf = k.add(f2_make(__uint_as_float(byte_as_float_bits(ta, byte)), __uint_as_float(byte_as_float_bits(tb, byte))),
                        f2_make(-8388608.0f, -8388608.0f));
    return f2_fma(cf, f2_make(0x1.010102p-8f, 0x1.010102p-8f), f2_mul(cf, f2_make(-0x1.fdfdfep-33f, -0x1.fdfdfep-33f)));
}
// taa.comp:32-38 on both lanes: y = (r + 2g) + b, co = 2r + (-2b), cg = (-r + 2g) + (-b)
VK_DEVICE void ycocg2(const Pk& k, f2 r, f2 g, f2 b, f2& y, f2& co, f2& cg)
{
    const f2 two = f2_make(2.0f, 2.0f);
    const f2 r2 = f2_mul(r, two), g2 = f2_mul(g, two), b2 = f2_mul(b, two);
    y = k.add(k.add(g2, r), b);
    co = k.sub(r2, b2);
    cg = k.sub(k.sub(g2, r), b);
}

struct TaaRow {           // one row of the window, this thread's two columns (a = lo lane, b = hi lane)
    float c[3][2];        // YCoCg of the two texels
    float hmn[3][2], hmx[3][2];   // min / max over the three texels (x-1, x, x+1) of the row
    uint2 bits;           // the two BGRA8 texels
};

VK_DEVICE float shfl_up1(float v, int lane)
{
#ifndef VKPBRT_HOSTSIM
    return __shfl_up_sync(0xffffffffu, v, 1);
#else
    return __shfl_sync(0xffffffffu, v, lane > 0 ? lane - 1 : 0);
#endif
}
VK_DEVICE float shfl_down1(float v, int lane)
{
#ifndef VKPBRT_HOSTSIM
    return __shfl_down_sync(0xffffffffu, v, 1);
#else
    return __shfl_sync(0xffffffffu, v, lane < 31 ? lane + 1 : 31);
#endif
}

struct TaaLane {          // per-thread constants of the strip
    int W, H, lane, xa, xl, y0, y1;
    bool dup_lo, dup_hi, writes;
};
struct TaaFetch {         // loads issued one step ahead (software pipeline: a strip is a chain of dependent rows)
    uint2 bits;           // denoised texels of the row the next step brings in
    uint2 mv;             // motion of the row the next step finishes
};

VK_DEVICE uint2 taa_load_bits(const TaaParams& p, const TaaLane& t, int r)
{
    const int ry = r < 0 ? 0 : (r >= t.H ? t.H - 1 : r);
    return __ldg(reinterpret_cast<const uint2*>(p.denoised + (size_t)ry * t.W + t.xl));           // BGRA8: byte0 = B
}
VK_DEVICE uint2 taa_load_mv(const TaaParams& p, const TaaLane& t, int gy)
{
    // rows outside the strip / lanes that do not write are never consumed: any valid address will do
    const int ry = gy < t.y0 ? t.y0 : (gy >= t.y1 ? t.y1 - 1 : gy);
    return __ldg(reinterpret_cast<const uint2*>(p.motion + (size_t)ry * t.W + t.xl));
}

// Step r brings row r into `bot` and, from the third row of the strip on, finishes the window's middle row (r - 1).
// Order inside a step: (a) request the NEXT step's texels and motion, (b) from this step's motion (requested one
// step ago) set up the bilinear taps and request the eight history texels, (c) convert row r to YCoCg and form its
// row minima / maxima while those are in flight, (d) test and blend.
VK_DEVICE void taa_step(const TaaParams& p, const Pk& k, const TaaLane& t, int r, const TaaRow& top, const TaaRow& mid, TaaRow& bot,
                        TaaFetch& f)
{
    const int W = t.W, H = t.H, lane = t.lane;
    if (r > t.y1) return;
    const bool finishes = (r >= t.y0 + 1) & t.writes;
    // (a)
    uint2 bits = f.bits;
    const uint2 mvb = f.mv;
    f.bits = taa_load_bits(p, t, r + 1);
    f.mv = taa_load_mv(p, t, r);
    // (b) :57-60 ('&': no short-circuit branches), :88 bilinear history of both pixels (common.cuh bilin_setup);
    // a pixel that is not reprojected samples uv = (0, 0) and keeps the current colour by a select
    const float u0 = f16_bits_to_f32((uint16_t)(mvb.x & 0xffffu)), v0 = f16_bits_to_f32((uint16_t)(mvb.x >> 16));
    const float u1 = f16_bits_to_f32((uint16_t)(mvb.y & 0xffffu)), v1 = f16_bits_to_f32((uint16_t)(mvb.y >> 16));
    const bool in0 = finishes & (p.frame != 0) & (u0 >= 0.0f) & (v0 >= 0.0f) & (u0 <= 1.0f) & (v0 <= 1.0f);
    const bool in1 = finishes & (p.frame != 0) & (u1 >= 0.0f) & (v1 >= 0.0f) & (u1 <= 1.0f) & (v1 <= 1.0f);
    const f2 us = f2_make(in0 ? u0 : 0.0f, in1 ? u1 : 0.0f), vs = f2_make(in0 ? v0 : 0.0f, in1 ? v1 : 0.0f);
    const f2 mh = f2_make(-0.5f, -0.5f), p1 = f2_make(1.0f, 1.0f);
    const f2 x = k.add(f2_mul(us, f2_dup((float)W)), mh), y = k.add(f2_mul(vs, f2_dup((float)H)), mh);
    const float fx0 = floorf(f2_lo(x)), fx1 = floorf(f2_hi(x)), fy0 = floorf(f2_lo(y)), fy1 = floorf(f2_hi(y));
    uint32_t h[2][4];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const int ix = (int)(l ? fx1 : fx0), iy = (int)(l ? fy1 : fy0);
        const int xA = ix < 0 ? ix + W : ix, xB = ix + 1 >= W ? ix + 1 - W : ix + 1;
        const int yA = iy < 0 ? iy + H : iy, yB = iy + 1 >= H ? iy + 1 - H : iy + 1;
        h[l][0] = __ldg(p.history + (uint32_t)(yA * W + xA));
        h[l][1] = __ldg(p.history + (uint32_t)(yA * W + xB));
        h[l][2] = __ldg(p.history + (uint32_t)(yB * W + xA));
        h[l][3] = __ldg(p.history + (uint32_t)(yB * W + xB));
    }
    // (c)
    bits = t.dup_lo ? make_uint2(bits.x, bits.x) : (t.dup_hi ? make_uint2(bits.y, bits.y) : bits);
    bot.bits = bits;
    {
        f2 yy, co, cg;
        ycocg2(k, unorm8_pair(k, bits.x, bits.y, 2), unorm8_pair(k, bits.x, bits.y, 1), unorm8_pair(k, bits.x, bits.y, 0), yy, co, cg);
        bot.c[0][0] = f2_lo(yy); bot.c[0][1] = f2_hi(yy);
        bot.c[1][0] = f2_lo(co); bot.c[1][1] = f2_hi(co);
        bot.c[2][0] = f2_lo(cg); bot.c[2][1] = f2_hi(cg);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float left = shfl_up1(bot.c[c][1], lane), right = shfl_down1(bot.c[c][0], lane);
            bot.hmn[c][0] = fminf(fminf(left, bot.c[c][0]), bot.c[c][1]);
            bot.hmx[c][0] = fmaxf(fmaxf(left, bot.c[c][0]), bot.c[c][1]);
            bot.hmn[c][1] = fminf(fminf(bot.c[c][0], bot.c[c][1]), right);
            bot.hmx[c][1] = fmaxf(fmaxf(bot.c[c][0], bot.c[c][1]), right);
        }
    }
    if (!finishes) return;
    // (d)
    const int gy = r - 1;                                         // the finished row
    const size_t pix = (size_t)gy * W + t.xa;
    uint2 out = make_uint2(mid.bits.x | 0xff000000u, mid.bits.y | 0xff000000u);     // unorm8 -> float -> unorm8 is the identity
    if (in0 | in1) {
        const f2 a = k.sub(x, f2_make(fx0, fx1)), bt = k.sub(y, f2_make(fy0, fy1));
        const f2 oma = k.sub(p1, a), omb = k.sub(p1, bt);
        const f2 w00 = f2_mul(oma, omb), w10 = f2_mul(a, omb), w01 = f2_mul(oma, bt), w11 = f2_mul(a, bt);
        // Bytes are B,G,R,A of the previous final; sampled as RGBA8 the .x channel reads byte 0 (App. C-4)
        f2 prev[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int kb = p.fix_swizzle ? (2 - c) : c;
            prev[c] = k.add(k.add(k.add(f2_mul(w00, unorm8_pair(k, h[0][0], h[1][0], kb)), f2_mul(w10, unorm8_pair(k, h[0][1], h[1][1], kb))),
                                  f2_mul(w01, unorm8_pair(k, h[0][2], h[1][2], kb))), f2_mul(w11, unorm8_pair(k, h[0][3], h[1][3], kb)));
        }
        f2 pyc[3];
        ycocg2(k, prev[0], prev[1], prev[2], pyc[0], pyc[1], pyc[2]);
        // :62-83 box (3x3) and cross (5 texel) bounds from the window, :90-96 the test
        bool inside0 = in0, inside1 = in1;
        const f2 hf = f2_make(0.5f, 0.5f);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float mnb[2], mxb[2], mnc[2], mxc[2];
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                mnb[l] = fminf(fminf(top.hmn[c][l], mid.hmn[c][l]), bot.hmn[c][l]);
                mxb[l] = fmaxf(fmaxf(top.hmx[c][l], mid.hmx[c][l]), bot.hmx[c][l]);
                mnc[l] = fminf(fminf(mid.hmn[c][l], top.c[c][l]), bot.c[c][l]);
                mxc[l] = fmaxf(fmaxf(mid.hmx[c][l], top.c[c][l]), bot.c[c][l]);
            }
            const f2 mn = f2_mul(k.add(f2_make(mnb[0], mnb[1]), f2_make(mnc[0], mnc[1])), hf);
            const f2 mx = f2_mul(k.add(f2_make(mxb[0], mxb[1]), f2_make(mxc[0], mxc[1])), hf);
            inside0 = inside0 & (f2_lo(pyc[c]) >= f2_lo(mn)) & (f2_lo(pyc[c]) <= f2_lo(mx));
            inside1 = inside1 & (f2_hi(pyc[c]) >= f2_hi(mn)) & (f2_hi(pyc[c]) <= f2_hi(mx));
        }
        // :98-103 res = .4 * cur + (1 - .4) * prev
        uint32_t o0 = 0xff000000u, o1 = 0xff000000u;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const f2 cur = unorm8_pair(k, mid.bits.x, mid.bits.y, 2 - c);
            const f2 res = k.add(f2_mul(f2_make(.4f, .4f), cur), f2_mul(f2_make(1 - .4f, 1 - .4f), prev[c]));
            // f32_to_unorm8 (common.cuh) without its NaN and lower clamps: res is a sum of products of values in [0, 1]
            // (unorm8 texels, bilinear weights) -- finite and >= 0 -- so only the upper clamp can act
            const f2 q = k.add(f2_mul(f2_make(fminf(f2_lo(res), 1.0f), fminf(f2_hi(res), 1.0f)), f2_make(255.0f, 255.0f)), f2_make(0.5f, 0.5f));
            o0 |= (uint32_t)f2_lo(q) << (8 * (2 - c));
            o1 |= (uint32_t)f2_hi(q) << (8 * (2 - c));
        }
        out.x = inside0 ? o0 : out.x;
        out.y = inside1 ? o1 : out.y;
    }
    *reinterpret_cast<uint2*>(p.final_bgra + pix) = out;
}

#ifndef TAA_MIN_CTAS
#define TAA_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(TAA_WARPS * 32, TAA_MIN_CTAS) k_taa(const TaaParams p)
{
    TaaLane t;
    t.W = p.W; t.H = p.H;
    t.lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int wx = blockIdx.x * TAA_WARPS + warp;
    if (wx * TAA_COLS >= t.W) return;                                     // whole warp
    t.xa = wx * TAA_COLS - 2 + 2 * t.lane;                                // this thread's columns xa, xa + 1 (W is even)
    // grid rows: the strips of [row_begin, row_end), then those of the optional second range (band-sharded runs: both edge
    // rows of a band in one launch)
    int strip = (int)blockIdx.y, rb = p.row_begin, re = p.row_end;
    if (strip >= p.strips) { strip -= p.strips; rb = p.row_begin2; re = p.row_end2; }
    t.y0 = rb + strip * p.rows_per_warp;
    t.y1 = (t.y0 + p.rows_per_warp < re) ? t.y0 + p.rows_per_warp : re;
    // Out-of-image neighbours are skipped by the shader (taa.comp:70); clamping the coordinate instead re-reads a texel
    // that is already in the same box / cross set, so min and max are unchanged.  A pair left of the image is texel 0
    // twice, a pair right of it texel W-1 twice.
    t.xl = t.xa < 0 ? 0 : (t.xa >= t.W ? t.W - 2 : t.xa);
    t.dup_lo = t.xa < 0; t.dup_hi = t.xa >= t.W;
    t.writes = (t.lane >= 1) & (t.lane <= 30) & (t.xa >= 0) & (t.xa < t.W);
    const Pk k{f2_dup(p.one), f2_dup(p.neg_one)};
    TaaFetch f;
    f.bits = taa_load_bits(p, t, t.y0 - 1);
    f.mv = taa_load_mv(p, t, t.y0);
    // rows y0-1 .. y1, one per step; the three window rows rotate through A, B, C by renaming (no register moves)
    TaaRow A, B, C;
    taa_step(p, k, t, t.y0 - 1, A, A, B, f);     // nothing is finished by the first two rows: the window arguments are unused
    taa_step(p, k, t, t.y0, A, B, C, f);
    for (int r = t.y0 + 1; r <= t.y1; r += 3) {
        taa_step(p, k, t, r, B, C, A, f);
        taa_step(p, k, t, r + 1, C, A, B, f);
        taa_step(p, k, t, r + 2, A, B, C, f);
    }
}

// one pixel per thread (first version): odd widths / unaligned planes only
constexpr int TAA_BX = 32, TAA_BY = 16;                     // output pixels per CTA

constexpr int TAA_TW = TAA_BX + 2, TAA_TH = TAA_BY + 2;     // YCoCg tile with a 1-texel apron

__global__ void __launch_bounds__(TAA_BX* TAA_BY) k_taa_scalar(const TaaParams p)
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

// strip height of the pair kernel.  A warp walks its strip row by row with dependent gathers (~1.4 us per row on B200), so
// a launch lasts at least strip-height steps: TAA_ROWS amortises the two lead-in rows when the grid fills the GPU several
// times over, but a band of a sharded frame (or a small image) would leave most warp slots empty for that long.  Then the
// strips shrink until the launch fills the resident warp slots about once.
static int taa_strip_rows(int rows, int warps_x)
{
    int sms = 148;
#ifndef VKPBRT_HOSTSIM
    static std::atomic<int> cached{0};
    sms = cached.load(std::memory_order_relaxed);
    if (sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached.store(sms = n, std::memory_order_relaxed);
    }
#endif
    const int slots = sms * TAA_MIN_CTAS * TAA_WARPS;                       // resident warps (register-limited)
    const int ctas_x = (warps_x + TAA_WARPS - 1) / TAA_WARPS;
    const int strips_that_fit = slots / (ctas_x * TAA_WARPS);              // per launch, one resident round
    if ((rows + TAA_ROWS - 1) / TAA_ROWS >= strips_that_fit || strips_that_fit <= 0) return TAA_ROWS;
    const int h = (rows + strips_that_fit - 1) / strips_that_fit;          // < TAA_ROWS
    return h < 1 ? 1 : h;
}

cudaError_t launch_taa(const TaaParams& p_in, cudaStream_t stream)
{
    TaaParams p = p_in;
    if (p.row_end < p.row_begin) p.row_end = p.row_begin;
    if (p.row_end2 < p.row_begin2) p.row_end2 = p.row_begin2;
    const int rows1 = p.row_end - p.row_begin, rows2 = p.row_end2 - p.row_begin2;
    if (rows1 + rows2 <= 0) return cudaSuccess;
    if (p.one != 1.0f || p.neg_one != -1.0f) return cudaErrorInvalidValue;
    const uintptr_t al = (uintptr_t)p.denoised | (uintptr_t)p.motion | (uintptr_t)p.final_bgra;
    if (p.W % 2 == 0 && p.W >= 2 && al % 8 == 0 && !p.force_scalar) {
        const int warps_x = (p.W + TAA_COLS - 1) / TAA_COLS;
        if (p.rows_per_warp <= 0) p.rows_per_warp = taa_strip_rows(rows1 + rows2, warps_x);      // > 0: the caller's (tests)
        p.strips = (rows1 + p.rows_per_warp - 1) / p.rows_per_warp;
        const int strips2 = (rows2 + p.rows_per_warp - 1) / p.rows_per_warp;
        dim3 block(TAA_WARPS * 32, 1, 1), grid((warps_x + TAA_WARPS - 1) / TAA_WARPS, p.strips + strips2, 1);
        VKPBRT_LAUNCH(k_taa, grid, block, 0, stream, p);
    } else {
        // one pixel per thread: one launch per range
        for (int part = 0; part < 2; ++part) {
            TaaParams q = p;
            if (part == 1) { q.row_begin = p.row_begin2; q.row_end = p.row_end2; }
            const int rows = q.row_end - q.row_begin;
            if (rows <= 0) continue;
            dim3 block(TAA_BX, TAA_BY, 1), grid((p.W + TAA_BX - 1) / TAA_BX, (rows + TAA_BY - 1) / TAA_BY, 1);
            VKPBRT_LAUNCH(k_taa_scalar, grid, block, 0, stream, q);
        }
    }
    return cudaGetLastError();
}

}  // namespace vkpbrt
