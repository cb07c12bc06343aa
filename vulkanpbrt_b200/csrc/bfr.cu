// bfr.cu -- k_bfr_block (shaders/bfr.comp:202-309) and k_bfr_blend (shaders/bfrBlender.comp:22-68).
//
// k_bfr_block: per jittered/mirrored BxB block, fit 7 features x 3 channels by 40 steps of the
// reference's Adam-style gradient descent (bfr.comp:100-142, :260-278), then the same temporal
// accumulation / tone-map epilogue as BMFR.  The reference runs B*B invocations (1 pixel each) and
// pays, per step, 22 subgroupAdds + a serial fold over up to 32 subgroups + 2 barriers.  Here a
// thread owns S pixels (T = 256 / 64 / 32 threads for B = 32 / 16 / 8) laid out so that a warp's 32
// lanes are exactly one of the reference's subgroups for each s; the 21 gradient terms + the L1
// count of a subgroup are reduced by ONE transposing shuffle network (23 shuffles) instead of 22
// butterflies, and 21 lanes of warp 0 own one (feature, channel) coefficient
// each, like bfr.comp's "ID < ALPHA_SIZE" invocations own one vec3.  For B = 8 the block is a
// single warp.  The sums follow the oracle's order exactly (xor-butterfly per subgroup, serial fold
// over subgroups, no FMA contraction), so the descent trajectory -- which contains a sign()
// non-linearity -- is bit-identical.  Latency / FP32-pipe bound, not HBM bound.
//
// k_bfr_blend: streaming (2r+1)^2 luminance-window standard deviation and 3-way blend; arithmetic
// is written with non-contracted IEEE ops so the BGRA8 output is bit-exact against the oracle.
#include "common.cuh"
#include "kernels.h"

namespace vkpbrt {

// bfr.comp:92
__constant__ int c_bfr_offsets[16][2] = {{-7, -11}, {-14, -8}, {-5, -12}, {-15, -1}, {-5, -9}, {-1, -4},
                                         {-14, -7}, {0, -13},  {-5, -1},  {-1, 0},   {-15, -2}, {-14, -10},
                                         {-1, -1},  {-6, -3},  {0, -8},   {-10, -4}};

// Transposing warp reduction of N values per lane (N need not be a power of two): at each of the five xor stages
// the lower half of the lanes keeps the first ceil(N/2) values and the upper half the rest (an odd count is padded
// with a zero), so 22 values cost 11 + 6 + 3 + 2 + 1 = 23 shuffles instead of 22 x 5.  Every total is summed along
// the xor-butterfly tree (lane^16, ^8, ^4, ^2, ^1): bit-identical to 22 separate subgroupAdds.
template <int N, int OFF>
struct ReduceN {
    static VK_DEVICE float run(const float* v, int lane)
    {
        if constexpr (OFF == 0) {
            return v[0];
        } else {
            constexpr int h = (N + 1) / 2;
            const bool up = (lane & OFF) != 0;
            float nv[h];
#pragma unroll
            for (int j = 0; j < h; ++j) {
                const float lo = v[j];
                const float hi = (j + h < N) ? v[j + h] : 0.0f;
                const float send = up ? lo : hi;
                const float keep = up ? hi : lo;
                nv[j] = add_rn(keep, __shfl_xor_sync(0xffffffffu, send, OFF));
            }
            return ReduceN<h, OFF / 2>::run(nv, lane);
        }
    }
};
// which of the 22 totals a lane ends up with (22 -> 11 -> 6 -> 3 -> 2 -> 1), or -1 for the lanes left with padding
VK_DEVICE int reduce22_slot(int lane)
{
    const int p3 = (lane & 1) + 2 * ((lane >> 1) & 1);          // position among the 3 values before the last two stages
    const int p1 = p3 + 3 * ((lane >> 2) & 1) + 6 * ((lane >> 3) & 1);
    return (p3 < 3 && p1 < 11) ? p1 + 11 * ((lane >> 4) & 1) : -1;
}

template <int NSG>
struct BfrShared {
    float red[22][NSG + 1];  // per-subgroup totals; +1: the fold reads one row per lane
    float alpha[24];         // alpha_vec[7] (vec3) flattened j*3+c
    float zmin[NSG], zmax[NSG];
    float zrange[2];
    int stop;
    uint32_t thr[256];       // tone-map thresholds (common.cuh: tonemap_code)
};

// occupancy targets (CTAs per SM) of the three instantiations; A/B-tested on B200 (tools/bfr_variants.sh)
#ifndef BFR_CTAS_B32
#define BFR_CTAS_B32 2
#endif
#ifndef BFR_CTAS_B16
#define BFR_CTAS_B16 9
#endif

template <int B, int T>
__global__ void __launch_bounds__(T, (T == 256 ? BFR_CTAS_B32 : (T == 64 ? BFR_CTAS_B16 : 32))) k_bfr_block(const BfrParams p)
{
    constexpr int N = B * B;
    constexpr int S = N / T;
    constexpr int NW = T / 32;
    constexpr int NSG = N / 32;         // subgroups of the reference workgroup
    __shared__ BfrShared<NSG> sm;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int bx = blockIdx.x, by = blockIdx.y;
    const int W = p.W, H = p.H;
    const uint32_t frame = p.frame;
    const int ox = c_bfr_offsets[frame & 15u][0], oy = c_bfr_offsets[frame & 15u][1];
    // a block with no pixel inside the image stores nothing (bfr.comp:293 drops its pixels): skip its 40 descent steps
    if (bx * B + ox >= p.W || bx * B + ox + B <= 0 || by * B + oy >= p.H || by * B + oy + B <= 0) return;

    // pixel q = t + s*T of the block (gl_LocalInvocationIndex = ly*B + lx): subgroup q / 32 = warp + s*NW,
    // lane q % 32 = lane
    float fy[S], fz[S], fn[S][3], noisy[S][3];
    bool l1[S], in_img[S];
    size_t pixs[S];
    const int lx = t % B;
    const float fx = sub_rn(div_rn((float)lx, sub_rn((float)B, 1.0f)), 0.5f);    // bfr.comp:247
    float zmin = 0.0f, zmax = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int q = t + s * T;
        const int ly = q / B;
        const int ax = bx * B + lx + ox, ay = by * B + ly + oy;                 // :209 (offsets are ADDED)
        const int ix = mirror(ax, W), iy = mirror(ay, H);
        in_img[s] = (ax == ix) && (ay == iy);
        const size_t pix = (size_t)iy * W + ix;
        pixs[s] = pix;
        const uint2 nz = __ldg(p.noisy + pix);
        noisy[s][0] = f16_bits_to_f32((uint16_t)(nz.x & 0xffffu));
        noisy[s][1] = f16_bits_to_f32((uint16_t)(nz.x >> 16));
        noisy[s][2] = f16_bits_to_f32((uint16_t)(nz.y & 0xffffu));
        const float z = __ldg(p.depth + pix);
        const float2 nrm = __ldg(p.normal + pix);
        float sth, cth, sph, cph;
        vk_sincos(nrm.x, sth, cth);
        vk_sincos(nrm.y, sph, cph);
        fn[s][0] = mul_rn(cph, sth);
        fn[s][1] = mul_rn(sph, sth);
        fn[s][2] = cth;
        fz[s] = z;
        fy[s] = sub_rn(div_rn((float)ly, sub_rn((float)B, 1.0f)), 0.5f);
        const float pixel_spp = mul_rn(unorm8_to_f32((uint32_t)__ldg(p.spp + pix)), 256.0f);   // :229
        l1[s] = pixel_spp >= 10.0f;                                              // SPP_THRESH
        zmin = s == 0 ? z : gl_min(z, zmin);
        zmax = s == 0 ? z : gl_max(z, zmax);
    }
    // ---- :241-246 depth normalisation over the block ---------------------------------------
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        zmin = gl_min(__shfl_xor_sync(0xffffffffu, zmin, off), zmin);
        zmax = gl_max(__shfl_xor_sync(0xffffffffu, zmax, off), zmax);
    }
    if (lane == 0) { sm.zmin[warp] = zmin; sm.zmax[warp] = zmax; }
    if (t < 24) sm.alpha[t] = 0.0f;                                              // :235-237
    for (int i = t; i < 256; i += T) sm.thr[i] = __ldg(c_tonemap_thr + i);
    if (t == 0) sm.stop = 0;
    __syncthreads();
    if (t == 0) {
        float a = sm.zmin[0], b = sm.zmax[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) { a = gl_min(sm.zmin[w], a); b = gl_max(sm.zmax[w], b); }
        sm.zrange[0] = a; sm.zrange[1] = b;
    }
    __syncthreads();
    zmin = sm.zrange[0];
    zmax = sm.zrange[1];
    const float zden = add_rn(sub_rn(zmax, zmin), 1e-8f);                        // EPS 1e-8 (:81, :244)
#pragma unroll
    for (int s = 0; s < S; ++s) fz[s] = sub_rn(mul_rn(div_rn(sub_rn(fz[s], zmin), zden), 2.0f), 1.0f);

    // ---- :260-278 gradient descent ---------------------------------------------------------
    float m = 0.0f, v = 0.0f;          // Adam moments of coefficient t (threads 0..20)
    const int slot = reduce22_slot(lane);
    float al[21];
    for (int it = 0; it < 40; ++it) {
#pragma unroll
        for (int j = 0; j < 21; ++j) al[j] = sm.alpha[j];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float f[7] = {1.0f, fx, fy[s], fz[s], fn[s][0], fn[s][1], fn[s][2]};
            float r[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float pred = 0.0f;
#pragma unroll
                for (int j = 0; j < 7; ++j) pred = add_rn(pred, mul_rn(f[j], al[j * 3 + c]));   // :262-264
                const float d = sub_rn(noisy[s][c], pred);
                const float sgn = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);          // sign(), :267-268
                r[c] = l1[s] ? sgn : d;                                                      // selects: no branch per pixel and step
            }
            float part[22];
#pragma unroll
            for (int j = 0; j < 7; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) part[j * 3 + c] = mul_rn(f[j], r[c]);                // :272-274
            part[21] = l1[s] ? 1.0f : 0.0f;
            const float tot = ReduceN<22, 16>::run(part, lane);                  // subgroupAdd x22 (:101-104)
            if (slot >= 0) sm.red[slot][warp + s * NW] = tot;
        }
        __syncthreads();
        if (t < 21) {
            // serial fold over the subgroups (the reference's order).  Every other thread of the block waits for this
            // section: for the 32 subgroups of a 32x32 block all partials are fetched up front, so only the dependent
            // additions remain in sequence (-5 % kernel time; no gain for 8 or 2 subgroups)
            float delta, cnt;
            if constexpr (NSG > 8) {
                float rd[NSG], rc[NSG];
#pragma unroll
                for (int g = 0; g < NSG; ++g) { rd[g] = sm.red[t][g]; rc[g] = sm.red[21][g]; }
                delta = rd[0];
                cnt = rc[0];
#pragma unroll
                for (int g = 1; g < NSG; ++g) { delta = add_rn(delta, rd[g]); cnt = add_rn(cnt, rc[g]); }
            } else {
                delta = sm.red[t][0];
                cnt = sm.red[21][0];
#pragma unroll
                for (int g = 1; g < NSG; ++g) { delta = add_rn(delta, sm.red[t][g]); cnt = add_rn(cnt, sm.red[21][g]); }
            }
            // :127-134
            const float l1_ratio = div_rn(mul_rn(cnt, 1.0f), (float)N);
            const float oml = sub_rn(1.0f, l1_ratio);
            const float a = add_rn(mul_rn(l1_ratio, 1.1f), mul_rn(oml, .863f));
            const float beta1 = add_rn(mul_rn(l1_ratio, .45f), mul_rn(oml, .3f));
            const float beta2 = add_rn(mul_rn(l1_ratio, .75f), mul_rn(oml, .7314f));
            m = add_rn(mul_rn(beta1, m), mul_rn((1.0f - .3f), delta));
            v = add_rn(mul_rn(beta2, v), mul_rn((1.0f - .7314f), mul_rn(fabsf(delta), fabsf(delta))));
            const float step = div_rn(mul_rn(mul_rn(a, p.lr_exp[it]), p.lr_sqrt[it]), p.lr_den[it]);
            sm.alpha[t] = add_rn(sm.alpha[t], mul_rn(step, div_rn(m, add_rn(sqrt_rn(v), 1e-8f))));
            // gradient_rest turns NaN -> the loop condition fails (:260).  gradient_left is the subgroupAdd of |delta| over
            // the threads with id < 7 that share thread id 0's subgroup (:136-137): with ID = x * size.y + y (:75) those are
            // the features j < 32 / B, so only their gradients can end the loop (oracle/vkpbrt_oracle.c, same place)
            if (t < 3 * (32 / B) && isnan(delta)) sm.stop = 1;
        }
        __syncthreads();
        if (sm.stop) break;
    }

    // ---- :282-308 ----------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < 21; ++j) al[j] = sm.alpha[j];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        if (!in_img[s]) continue;
        const float f[7] = {1.0f, fx, fy[s], fz[s], fn[s][0], fn[s][1], fn[s][2]};
        float cr = 0.0f, cg = 0.0f, cb = 0.0f;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            cr = add_rn(cr, mul_rn(f[j], al[j * 3 + 0]));
            cg = add_rn(cg, mul_rn(f[j], al[j * 3 + 1]));
            cb = add_rn(cb, mul_rn(f[j], al[j * 3 + 2]));
        }
        cr = gl_clamp(cr, 0.0f, 10.0f);
        cg = gl_clamp(cg, 0.0f, 10.0f);
        cb = gl_clamp(cb, 0.0f, 10.0f);
        const size_t pix = pixs[s];
        denoise_epilogue(cr, cg, cb, frame, pix, W, H, __ldg(p.motion + pix), (uint32_t)__ldg(p.spp + pix),
                         __ldg(p.albedo + pix), p.denoised_prev, p.denoised_next, p.final_bgra, sm.thr);
    }
}

cudaError_t launch_bfr(const BfrParams& p, cudaStream_t stream)
{
    dim3 grid(p.blocks_x, p.blocks_y, 1);
    if (p.block == 32) {
        VKPBRT_LAUNCH((k_bfr_block<32, 256>), grid, dim3(256, 1, 1), 0, stream, p);
    } else if (p.block == 16) {
        VKPBRT_LAUNCH((k_bfr_block<16, 64>), grid, dim3(64, 1, 1), 0, stream, p);
    } else if (p.block == 8) {
        VKPBRT_LAUNCH((k_bfr_block<8, 32>), grid, dim3(32, 1, 1), 0, stream, p);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// bfrBlender.comp
// ------------------------------------------------------------------------------------------------
VK_DEVICE float mix3(float a, float b, float c, float t, float mid, float max_dev)
{
    t = __fdiv_rn(t, max_dev);
    t = gl_min(t, 1.0f);
    const float a_fac = gl_max(sub_rn(1.0f, __fdiv_rn(t, mid)), 0.0f);
    const float b_fac = (t < mid) ? __fdiv_rn(t, mid) : sub_rn(1.0f, __fdiv_rn(sub_rn(t, mid), sub_rn(1.0f, mid)));
    const float c_fac = sub_rn(sub_rn(1.0f, a_fac), b_fac);
    return add_rn(add_rn(mul_rn(a_fac, a), mul_rn(b_fac, b)), mul_rn(c_fac, c));
}

VK_DEVICE float lum3(float r, float g, float b)
{
    const float third = 0.3333333432674408f;   // float(1.0 / 3.0)
    return add_rn(add_rn(mul_rn(r, third), mul_rn(g, third)), mul_rn(b, third));
}

// Round 2: the window statistics from a shared LUMINANCE tile.  The first version (k_bfr_blend_generic below, still
// used for radii other than 1..3) gathered the (2r+1)^2 rgba16f texels per pixel from global memory, converted and
// averaged each of them again for every pixel that touches it and divided by the running count with an IEEE division
// per tap: 1500 instructions per pixel, 10 % of the HBM roofline.  Here a CTA stages the luminance of its 64 x 8 pixels
// plus an r-texel apron ONCE per texel, every thread folds the windows of two pixels (x and x + 32) at a time in packed
// fp32 pairs, and the weights 1/n and 1 - 1/n are compile-time constants (same bits as the division: correctly rounded
// either way).  mix3's divisions by mid = 0.5, 1 - mid = 0.5 and max_dev = 1.0 are exact scalings, written as such.
template <int R>
struct BlendWeights {           // w[n-1] = RN(1 / n), omw[n-1] = RN(1 - w), n = 1 .. (2R+1)^2
    float w[(2 * R + 1) * (2 * R + 1)], omw[(2 * R + 1) * (2 * R + 1)];
    constexpr BlendWeights() : w{}, omw{}
    {
        for (int n = 1; n <= (2 * R + 1) * (2 * R + 1); ++n) {
            w[n - 1] = 1.0f / (float)n;
            omw[n - 1] = 1.0f - w[n - 1];
        }
    }
};

VK_DEVICE float mix3_half(float a, float b, float c, float t)      // mix3(a, b, c, t, mid = .5, max_dev = 1)
{
    t = gl_min(t, 1.0f);                                            // t / 1.0f == t
    const float t2 = mul_rn(t, 2.0f);                               // t / 0.5f, exact
    const float a_fac = gl_max(sub_rn(1.0f, t2), 0.0f);
    const float b_fac = (t < .5f) ? t2 : sub_rn(1.0f, mul_rn(sub_rn(t, .5f), 2.0f));
    const float c_fac = sub_rn(sub_rn(1.0f, a_fac), b_fac);
    return add_rn(add_rn(mul_rn(a_fac, a), mul_rn(b_fac, b)), mul_rn(c_fac, c));
}

template <int R>
__global__ void __launch_bounds__(256) k_bfr_blend(const BlendParams p)
{
    constexpr int TW = 64 + 2 * R, TH = 8 + 2 * R;
    __shared__ float lum[TH][TW];
    const int W = p.W, H = p.H;
    const int x0 = blockIdx.x * 64, y0 = blockIdx.y * 8;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < TW * TH; i += 256) {
        const int ty = i / TW, tx = i - ty * TW;
        const int sx = x0 + tx - R, sy = y0 + ty - R;
        float cr = 0.0f, cg = 0.0f, cb = 0.0f;                                  // out of range: robust-access 0
        if (sx >= 0 && sy >= 0 && sx < W && sy < H) load_rgb16f(p.average, (size_t)sy * W + sx, cr, cg, cb);
        lum[ty][tx] = lum3(cr, cg, cb);
    }
    __syncthreads();
    const int gx = x0 + threadIdx.x, gy = y0 + threadIdx.y;                     // pixels (gx, gy) and (gx + 32, gy)
    if (gy >= H || gx >= W) return;                                             // bfrBlender.comp:32
    const bool has1 = gx + 32 < W;
    const Pk k{f2_dup(p.one), f2_dup(p.neg_one)};
    constexpr BlendWeights<R> wt{};
    f2 sq = f2_make(0.0f, 0.0f), av = sq;
#pragma unroll
    for (int dy = 0; dy <= 2 * R; ++dy)
#pragma unroll
        for (int dx = 0; dx <= 2 * R; ++dx) {
            constexpr int dummy = 0; (void)dummy;
            const int n = dy * (2 * R + 1) + dx;                                // taps in the shader's order: y outer, x inner
            const f2 cur = f2_make(lum[threadIdx.y + dy][threadIdx.x + dx], lum[threadIdx.y + dy][threadIdx.x + 32 + dx]);
            const f2 w = f2_dup(wt.w[n]), omw = f2_dup(wt.omw[n]);
            sq = k.add(f2_mul(sq, omw), f2_mul(f2_mul(cur, cur), w));           // :42-43 mix(x, y, 1/n) = x * (1 - 1/n) + y * (1/n)
            av = k.add(f2_mul(av, omw), f2_mul(cur, w));
        }
    const size_t pix0 = (size_t)gy * W + gx, pix1 = has1 ? pix0 + 32 : pix0;
    float ar0, ag0, ab0, qr0, qg0, qb0, ar1, ag1, ab1, qr1, qg1, qb1;
    load_rgb16f(p.average_squared, pix0, qr0, qg0, qb0);
    load_rgb16f(p.average_squared, pix1, qr1, qg1, qb1);
    load_rgb16f(p.average, pix0, ar0, ag0, ab0);
    load_rgb16f(p.average, pix1, ar1, ag1, ab1);
    const f2 hf = f2_make(.5f, .5f);
    av = k.add(f2_mul(av, hf), f2_mul(f2_make(lum3(ar0, ag0, ab0), lum3(ar1, ag1, ab1)), hf));       // :46-49 (1 - .5 == .5)
    sq = k.add(f2_mul(sq, hf), f2_mul(f2_make(lum3(qr0, qg0, qb0), lum3(qr1, qg1, qb1)), hf));
    const f2 var = k.sub(sq, f2_mul(av, av));
    const float sd[2] = {__fsqrt_rn(f2_lo(var)), __fsqrt_rn(f2_hi(var))};       // :58 (NaN for a negative argument, as in the shader)
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        if (l == 1 && !has1) break;
        const size_t pix = l ? pix1 : pix0;
        const uint32_t d0 = __ldg(p.denoised0 + pix), d1 = __ldg(p.denoised1 + pix), d2 = __ldg(p.denoised2 + pix);
        uint32_t out = 0xff000000u;
#pragma unroll
        for (int c = 0; c < 3; ++c) {      // c indexes BGRA8 memory bytes; the blend is per channel
            const float den2 = unorm8_byte_to_f32(d0, c);                       // :50-52 binding swap
            const float den1 = unorm8_byte_to_f32(d1, c);
            const float den0 = unorm8_byte_to_f32(d2, c);
            out |= (uint32_t)f32_to_unorm8(mix3_half(den0, den1, den2, sd[l])) << (8 * c);
        }
        p.final_bgra[pix] = out;
    }
}

// any radius: one pixel per thread, global gathers (first version)
__global__ void __launch_bounds__(256) k_bfr_blend_generic(const BlendParams p)
{
    const int gx = blockIdx.x * 32 + threadIdx.x, gy = blockIdx.y * 8 + threadIdx.y;
    if (gx >= p.W || gy >= p.H) return;                                         // bfrBlender.comp:32
    const int W = p.W, H = p.H, r = p.radius;
    float sq = 0.0f, av = 0.0f;
    int count = 0;
    for (int y = -r; y <= r; ++y)
        for (int x = -r; x <= r; ++x) {
            const int sx = gx + x, sy = gy + y;
            float cr = 0.0f, cg = 0.0f, cb = 0.0f;                              // out of range: robust-access 0
            if (sx >= 0 && sy >= 0 && sx < W && sy < H) load_rgb16f(p.average, (size_t)sy * W + sx, cr, cg, cb);
            const float cur = lum3(cr, cg, cb);
            ++count;
            const float w = __fdiv_rn(1.0f, (float)count);
            sq = gl_mix_exact(sq, mul_rn(cur, cur), w);                         // :42-43
            av = gl_mix_exact(av, cur, w);
        }
    const size_t pix = (size_t)gy * W + gx;
    float ar, ag, ab, qr, qg, qb;
    load_rgb16f(p.average, pix, ar, ag, ab);
    load_rgb16f(p.average_squared, pix, qr, qg, qb);
    av = gl_mix_exact(av, lum3(ar, ag, ab), .5f);                               // :46-49
    sq = gl_mix_exact(sq, lum3(qr, qg, qb), .5f);
    const float std_dev = __fsqrt_rn(sub_rn(sq, mul_rn(av, av)));               // :58
    const uint32_t d0 = __ldg(p.denoised0 + pix), d1 = __ldg(p.denoised1 + pix), d2 = __ldg(p.denoised2 + pix);
    uint32_t out = 0xff000000u;
#pragma unroll
    for (int c = 0; c < 3; ++c) {      // c indexes BGRA8 memory bytes; the blend is per channel
        const float den2 = unorm8_byte_to_f32(d0, c);                           // :50-52 binding swap
        const float den1 = unorm8_byte_to_f32(d1, c);
        const float den0 = unorm8_byte_to_f32(d2, c);
        out |= (uint32_t)f32_to_unorm8(mix3(den0, den1, den2, std_dev, .5f, 1.0f)) << (8 * c);
    }
    p.final_bgra[pix] = out;
}

cudaError_t launch_bfr_blend(const BlendParams& p, cudaStream_t stream)
{
    if (p.one != 1.0f || p.neg_one != -1.0f) return cudaErrorInvalidValue;
    if (p.radius >= 1 && p.radius <= 3) {
        dim3 block(32, 8, 1), grid((p.W + 63) / 64, (p.H + 7) / 8, 1);
        if (p.radius == 1) { VKPBRT_LAUNCH((k_bfr_blend<1>), grid, block, 0, stream, p); }
        else if (p.radius == 2) { VKPBRT_LAUNCH((k_bfr_blend<2>), grid, block, 0, stream, p); }
        else { VKPBRT_LAUNCH((k_bfr_blend<3>), grid, block, 0, stream, p); }
    } else {
        dim3 block(32, 8, 1), grid((p.W + 31) / 32, (p.H + 7) / 8, 1);
        VKPBRT_LAUNCH(k_bfr_blend_generic, grid, block, 0, stream, p);
    }
    return cudaGetLastError();
}

}  // namespace vkpbrt
