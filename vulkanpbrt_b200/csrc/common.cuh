// common.cuh -- device helpers shared by the sm_100a denoising kernels.
//
// Storage conventions (DESIGN.md "Data layout in HBM"): every plane is pitch-linear and tightly
// packed in the reference's own texel formats, so fusing stages does not move any rounding point
// (SURVEY.md section 0 item 5 / App. C):
//   fp32 -> fp16 stores   round-to-nearest-even (cvt.rn.f16.f32)
//   fp32 -> unorm8 stores NaN -> 0, clamp, (uint8)(c*255 + 0.5)
//   texture()             bilinear, REPEAT, exact fp32 weights (App. A.1)
// The *_rn helpers below use __fmul_rn/__fadd_rn so the compiler cannot contract them into
// FMAs: they feed integer decisions (validity masks, unorm8 codes) that must be bit-exact.
#pragma once

#ifdef VKPBRT_HOSTSIM
#include "hostsim.h"   // tests/hostsim: CPU emulator used by the GPU-less debug tests only
#else
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#ifndef VKPBRT_HOSTSIM
#define VK_DEVICE __device__ __forceinline__
#define VKPBRT_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define VKPBRT_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#define VKPBRT_DYN_SMEM(name) unsigned char* name = hostsim::dyn_smem()
#define VK_DEVICE inline
#endif

namespace vkpbrt {

// GLSL min/max/clamp: NaN behaviour follows the ternary definitions of the GLSL spec, which is
// what the reference shaders get (fminf/fmaxf would drop NaNs).
VK_DEVICE float gl_min(float x, float y) { return (y < x) ? y : x; }
VK_DEVICE float gl_max(float x, float y) { return (x < y) ? y : x; }
VK_DEVICE float gl_clamp(float x, float lo, float hi) { return gl_min(gl_max(x, lo), hi); }

VK_DEVICE float mul_rn(float a, float b) { return __fmul_rn(a, b); }
VK_DEVICE float add_rn(float a, float b) { return __fadd_rn(a, b); }
VK_DEVICE float sub_rn(float a, float b) { return __fadd_rn(a, -b); }

VK_DEVICE float div_rn(float a, float b) { return __fdiv_rn(a, b); }
VK_DEVICE float sqrt_rn(float a) { return __fsqrt_rn(a); }

// a / b == RN(a / b) from a correctly rounded reciprocal rb = RN(1 / b) (Markstein): q0 = RN(a*rb),
// e = a - b*q0 (exact in one FMA), q = RN(q0 + e*rb).  Bit-identical to IEEE division whenever no
// intermediate leaves the normal range; callers guarantee that with the range predicates below
// (2^-40 <= |b| <= 2^40, |a| == 0 or 2^-84 <= |a| <= 2^84: the quotient stays normal and the residual e is a multiple of
// 2^-149; validated against a / b on 4e8 random and adversarial operand pairs, tools/markstein_check.c -- the same
// program reports mismatches from 2^90 on).  3 instructions instead of the ~12 + branch of
// the generic IEEE sequence.
VK_DEVICE float div_by_rcp(float a, float b, float rb)
{
    const float q0 = mul_rn(a, rb);
    const float e = fmaf(-q0, b, a);
    return fmaf(e, rb, q0);
}
// out-of-line generic division for the rare operands outside div_by_rcp's range (keeps the ~12-instruction
// IEEE sequence and its slow-path call out of the unrolled hot loops)
#ifndef VKPBRT_HOSTSIM
static __device__ __noinline__ float div_rn_cold(float a, float b) { return __fdiv_rn(a, b); }
#else
inline float div_rn_cold(float a, float b) { return a / b; }
#endif

VK_DEVICE bool in_pow2_range(float x, uint32_t lo_bits, uint32_t hi_bits)   // lo <= |x| <= hi (normal numbers)
{
    return ((__float_as_uint(x) & 0x7fffffffu) - lo_bits) <= (hi_bits - lo_bits);
}
VK_DEVICE bool safe_divisor(float b) { return in_pow2_range(b, 0x2b800000u, 0x53800000u); }        // 2^-40 .. 2^40
// One factor of a numerator that is the product of two: 0 or 2^-42 .. 2^42, so the product lies in div_by_rcp's range.
// (The first version used 2^-30: chance cancellations -- a dot product or a column entry of some 1e-10 in a block whose columns
// are at the 1e-4 noise level -- sent ~8 blocks of a 4K frame through the slow generic fit, which is what a band's
// kernel then waits for when such a block runs in its last wave.)
VK_DEVICE bool safe_factor(float x)
{
    // '|' on purpose: a short-circuit '||' (and '&&' chains of these tests in the callers) compiles to a branch inside a
    // convergence-barrier region per test, which costs far more in the unrolled Householder stream than the test itself
    const uint32_t a = __float_as_uint(x) & 0x7fffffffu;
    return (a == 0u) | ((a - 0x2a800000u) <= (0x54800000u - 0x2a800000u));
}
// a / b for a divisor whose reciprocal is reused: exact fast path when both operands are in range
VK_DEVICE float div_guarded(float a, float b, float rb, bool b_safe)
{
    const uint32_t ia = __float_as_uint(a) & 0x7fffffffu;
    const bool ok = b_safe & ((ia == 0u) | ((ia - 0x15800000u) <= (0x69800000u - 0x15800000u)));     // 0 or 2^-84 .. 2^84; one branch
    if (ok) return div_by_rcp(a, b, rb);
    return div_rn_cold(a, b);
}

// ---- packed fp32 pairs (sm_100: FMUL2 / FFMA2 issue two IEEE binary32 operations per instruction) ---------------
// The block kernels are bound by instruction ISSUE, not by the FP32 pipe, so pairing independent operations halves
// the slots their arithmetic needs while every lane still performs the same correctly rounded operation.
// Two rules keep the results bit-identical to the scalar sequence:
//   * ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 (single rounding) even with --fmad=false
//     and even when the addition is written as fma(x, 1.0, y) with a literal 1.0, so an addition that consumes a
//     product is issued as f2_fma(product, ONE, addend) with ONE = (1.0f, 1.0f) taken from a KERNEL PARAMETER:
//     RN(product * 1 + addend) is the addition, and the assembler cannot see the multiplier;
//   * a - b is f2_fma(b, NEG_ONE, a) for the same reason (no separate negation instruction either).
#ifndef VKPBRT_HOSTSIM
struct f2 { unsigned long long v; };
VK_DEVICE f2 f2_make(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
VK_DEVICE float f2_lo(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); (void)hi; return lo; }
VK_DEVICE float f2_hi(f2 a) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); (void)lo; return hi; }
VK_DEVICE f2 f2_mul(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
VK_DEVICE f2 f2_fma(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
#else
struct f2 { float lo, hi; };
inline f2 f2_make(float lo, float hi) { return f2{lo, hi}; }
inline float f2_lo(f2 a) { return a.lo; }
inline float f2_hi(f2 a) { return a.hi; }
inline f2 f2_mul(f2 a, f2 b) { return f2{a.lo * b.lo, a.hi * b.hi}; }
inline f2 f2_fma(f2 a, f2 b, f2 c) { return f2{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)}; }
#endif
VK_DEVICE f2 f2_dup(float x) { return f2_make(x, x); }
// arithmetic context of a kernel that works on pairs: the run-time 1.0 / -1.0
struct Pk {
    f2 one, neg_one;
    VK_DEVICE f2 add(f2 a, f2 b) const { return f2_fma(a, one, b); }              // RN(a + b)
    VK_DEVICE f2 sub(f2 a, f2 b) const { return f2_fma(b, neg_one, a); }          // RN(a - b)
    VK_DEVICE f2 neg(f2 a) const { return f2_mul(a, neg_one); }                   // exact
    // a / b from rb = RN(1 / b): div_by_rcp on both lanes
    VK_DEVICE f2 div_by_rcp(f2 a, f2 b, f2 rb) const
    {
        const f2 q0 = f2_mul(a, rb);
        const f2 e = f2_fma(q0, neg(b), a);
        return f2_fma(e, rb, q0);
    }
};
VK_DEVICE f2 f2_select(bool c, f2 a, f2 b) { return f2_make(c ? f2_lo(a) : f2_lo(b), c ? f2_hi(a) : f2_hi(b)); }

// ---- deterministic sin / cos / pow ------------------------------------------------------------
// GLSL leaves their precision to the implementation; the oracle (oracle/vkpbrt_oracle.c: vk_sincos,
// vk_pow) fixes them as plain-IEEE Cephes-style algorithms and these are the same operations in the
// same order, so normals, features and tone-mapped texels are bit-identical on CPU and GPU.
VK_DEVICE void vk_sincos(float x, float& sn, float& cs)
{
    // branch-free form of the oracle's vk_sincos: out-of-range arguments are evaluated on a dummy and replaced by
    // the oracle's NaN at the end; "if (j & 1) { j += 1; y += 1; }" is j = (j + 1) & ~1 with y = float(j) (exact, j < 2^24)
    const float ax0 = fabsf(x);
    const bool in_range = ax0 < 8192.0f;
    const float ax = in_range ? ax0 : 0.0f;
    const uint32_t j = ((uint32_t)mul_rn(ax, 1.27323954473516f) + 1u) & ~1u;
    const float y = (float)j;
    const float r = sub_rn(sub_rn(sub_rn(ax, mul_rn(y, 0.78515625f)), mul_rn(y, 2.4187564849853515625e-4f)), mul_rn(y, 3.77489497744594108e-8f));
    const float z = mul_rn(r, r);
    const float ps = add_rn(mul_rn(mul_rn(add_rn(mul_rn(add_rn(mul_rn(-1.9515295891e-4f, z), 8.3321608736e-3f), z), -1.6666654611e-1f), z), r), r);
    float pc = mul_rn(mul_rn(add_rn(mul_rn(add_rn(mul_rn(2.443315711809948e-5f, z), -1.388731625493765e-3f), z), 4.166664568298827e-2f), z), z);
    pc = sub_rn(pc, mul_rn(0.5f, z));
    pc = add_rn(pc, 1.0f);
    // quadrant q: (s, c) = (ps, pc), (pc, -ps), (-ps, -pc), (-pc, ps).  One select per output and the signs as xors of the
    // sign bit (exactly the unary minus): the four-way select chains compiled to three-way branches per call
    const uint32_t q = (j >> 1) & 3u;
    const bool odd = (q & 1u) != 0u;
    const uint32_t s_sign = ((q & 2u) << 30) ^ (x < 0.0f ? 0x80000000u : 0u);       // sin: q = 2, 3; odd function
    const uint32_t c_sign = ((q + 1u) & 2u) << 30;                                  // cos: q = 1, 2
    const float s = __uint_as_float(__float_as_uint(odd ? pc : ps) ^ s_sign);
    const float c = __uint_as_float(__float_as_uint(odd ? ps : pc) ^ c_sign);
    const float nan = __uint_as_float(0x7fc00000u);
    sn = in_range ? s : nan;
    cs = in_range ? c : nan;
}

VK_DEVICE float vk_pow(float x0, float y)
{
    // branch-free form of the oracle's vk_pow (every early exit becomes a select on the final value)
    const bool positive = x0 > 0.0f, huge = x0 > 3.0e38f;
    float x = (positive && !huge) ? x0 : 1.0f;
    const bool sub = x < 1.17549435e-38f;
    x = sub ? mul_rn(x, 8388608.0f) : x;
    const uint32_t bits = __float_as_uint(x);
    int e = (sub ? -23 : 0) + (int)((bits >> 23) & 255u) - 127;
    float m = __uint_as_float((bits & 0x007fffffu) | 0x3f800000u);
    const bool big = m > 1.41421356f;
    m = big ? mul_rn(m, 0.5f) : m;
    e += big ? 1 : 0;
    const float z = div_rn(sub_rn(m, 1.0f), add_rn(m, 1.0f));
    const float z2 = mul_rn(z, z);
    const float p = mul_rn(add_rn(mul_rn(add_rn(mul_rn(add_rn(mul_rn(add_rn(mul_rn(0.0909090909f, z2), 0.1111111111f), z2), 0.1428571429f), z2), 0.2f), z2), 0.3333333333f), z2);
    const float lnm = mul_rn(2.0f, add_rn(z, mul_rn(z, p)));
    const float lg = add_rn((float)e, mul_rn(lnm, 1.44269504089f));
    const float t0 = mul_rn(y, lg);
    const bool over = t0 > 127.99f, under = t0 < -150.0f;
    const float t = (over || under) ? 0.0f : t0;
    const float n = rintf(t);
    const float f = sub_rn(t, n);
    const float px = add_rn(mul_rn(add_rn(mul_rn(add_rn(mul_rn(add_rn(mul_rn(add_rn(mul_rn(1.535336188319500e-4f, f), 1.339887440266574e-3f), f),
                                                                    9.618437357674640e-3f), f), 5.550332471162809e-2f), f), 2.402264791363012e-1f), f), 6.931472028550421e-1f);
    const float r = add_rn(1.0f, mul_rn(f, px));
    const int ni = (int)n, n1 = ni / 2, n2 = ni - n1;
    float res = mul_rn(mul_rn(r, __uint_as_float((uint32_t)(n1 + 127) << 23)), __uint_as_float((uint32_t)(n2 + 127) << 23));
    res = over ? __uint_as_float(0x7f800000u) : (under ? 0.0f : res);
    res = huge ? x0 : res;
    // x <= 0 or NaN: pow(0, y) = 0, otherwise the oracle's (x - x) / (x - x) = NaN
    return positive ? res : ((x0 == 0.0f) ? 0.0f : __uint_as_float(0x7fc00000u));
}

// mix(x, y, a) = x*(1-a) + y*a, no contraction
VK_DEVICE float gl_mix_exact(float x, float y, float a) { return add_rn(mul_rn(x, sub_rn(1.0f, a)), mul_rn(y, a)); }

VK_DEVICE uint16_t f32_to_f16_bits(float f) { return __half_as_ushort(__float2half_rn(f)); }
VK_DEVICE float f16_bits_to_f32(uint16_t h) { return __half2float(__ushort_as_half(h)); }

VK_DEVICE uint8_t f32_to_unorm8(float c)
{
    c = (c == c) ? c : 0.0f;          // NaN -> 0
    c = c < 0.0f ? 0.0f : c;
    c = c > 1.0f ? 1.0f : c;
    return (uint8_t)add_rn(mul_rn(c, 255.0f), 0.5f);
}
// c / 255.0f, exactly rounded for every code 0..255 (checked by tests/test_oracle_kat.py against the oracle's
// division): with 1/255 = RH + RL as a double-float, c*RH + (c*RL) rounds to the quotient -- two operations
VK_DEVICE float unorm8_scale(float cf)
{
    return fmaf(cf, 0x1.010102p-8f, mul_rn(cf, -0x1.fdfdfep-33f));
}
VK_DEVICE float unorm8_to_f32(uint32_t c) { return unorm8_scale((float)c); }
// byte K of a packed texel: one PRMT builds 2^23 + byte as a float, one subtraction makes it the integer value
VK_DEVICE float unorm8_byte_to_f32(uint32_t texel, int k)
{
#ifndef VKPBRT_HOSTSIM
    const uint32_t bits = __byte_perm(texel, 0x4b000000u, 0x7650u + (uint32_t)k);
#else
    const uint32_t bits = 0x4b000000u | ((texel >> (8 * k)) & 0xffu);
#endif
    return unorm8_scale(sub_rn(__uint_as_float(bits), 8388608.0f));
}

// bmfrGeneral.comp:93-97 / bfr.comp:192-196
VK_DEVICE int mirror(int x, int s)
{
    const int below = -x - 1, above = 2 * s - x - 1;        // both candidates, then selects: no branch per pixel
    return x < 0 ? below : (x >= s ? above : x);
}

// ---- bilinear sampler, REPEAT addressing, normalised coordinates ----
struct Bilin {
    int x0, x1, y0, y1;
    float w00, w10, w01, w11;
};

VK_DEVICE int wrapi(int i, int n)
{
    i %= n;
    return i < 0 ? i + n : i;
}

VK_DEVICE Bilin bilin_setup(float u, float v, int W, int H)
{
    Bilin b;
    float x = sub_rn(mul_rn(u, (float)W), 0.5f);
    float y = sub_rn(mul_rn(v, (float)H), 0.5f);
    float fx0 = floorf(x), fy0 = floorf(y);
    float a = sub_rn(x, fx0), bt = sub_rn(y, fy0);
    int ix = (int)fx0, iy = (int)fy0;
    // callers guarantee uv in [0,1] so ix in [-1, W-1]: cheap wrap
    b.x0 = ix < 0 ? ix + W : ix;
    b.x1 = ix + 1 >= W ? ix + 1 - W : ix + 1;
    b.y0 = iy < 0 ? iy + H : iy;
    b.y1 = iy + 1 >= H ? iy + 1 - H : iy + 1;
    float oma = sub_rn(1.0f, a), omb = sub_rn(1.0f, bt);
    b.w00 = mul_rn(oma, omb);
    b.w10 = mul_rn(a, omb);
    b.w01 = mul_rn(oma, bt);
    b.w11 = mul_rn(a, bt);
    return b;
}

// bilin_setup for coordinates that did not come from k_accumulate (a caller-supplied motion plane, stale rows): the
// reference samples with a REPEAT sampler, which is defined for ANY uv.  Same weights; the integer taps are wrapped with
// a true modulo when they fall outside the cheap wrap's range, and a non-finite coordinate samples texel (0, 0) with
// NaN weights (what a sampler returns for NaN coordinates is undefined) instead of reading out of bounds.
VK_DEVICE Bilin bilin_setup_repeat(float u, float v, int W, int H)
{
    Bilin b = bilin_setup(u, v, W, H);
    const bool cheap_ok = ((unsigned)b.x0 < (unsigned)W) & ((unsigned)b.x1 < (unsigned)W) & ((unsigned)b.y0 < (unsigned)H) & ((unsigned)b.y1 < (unsigned)H);
    if (!cheap_ok) {
        const float x = sub_rn(mul_rn(u, (float)W), 0.5f), y = sub_rn(mul_rn(v, (float)H), 0.5f);
        const bool finite = (fabsf(x) < 1.0e9f) & (fabsf(y) < 1.0e9f);
        const int ix = finite ? (int)floorf(x) : 0, iy = finite ? (int)floorf(y) : 0;
        b.x0 = wrapi(ix, W); b.x1 = wrapi(ix + 1, W);
        b.y0 = wrapi(iy, H); b.y1 = wrapi(iy + 1, H);
    }
    return b;
}

VK_DEVICE float bilin_mix(const Bilin& b, float t00, float t10, float t01, float t11)
{
    return add_rn(add_rn(add_rn(mul_rn(b.w00, t00), mul_rn(b.w10, t10)), mul_rn(b.w01, t01)), mul_rn(b.w11, t11));
}

// rgba16f texel = 8 bytes
struct Half4 {
    uint16_t x, y, z, w;
};

VK_DEVICE void load_rgb16f(const uint2* __restrict__ plane, size_t idx, float& r, float& g, float& b)
{
    uint2 t = __ldg(plane + idx);
    r = f16_bits_to_f32((uint16_t)(t.x & 0xffffu));
    g = f16_bits_to_f32((uint16_t)(t.x >> 16));
    b = f16_bits_to_f32((uint16_t)(t.y & 0xffffu));
}

VK_DEVICE uint2 pack_rgba16f(float r, float g, float b, float a)
{
    uint2 t;
    t.x = (uint32_t)f32_to_f16_bits(r) | ((uint32_t)f32_to_f16_bits(g) << 16);
    t.y = (uint32_t)f32_to_f16_bits(b) | ((uint32_t)f32_to_f16_bits(a) << 16);
    return t;
}

// bilinear fetch of the rgb channels of an rgba16f plane
VK_DEVICE void sample_rgb16f(const uint2* __restrict__ plane, const Bilin& bl, int W, float& r, float& g, float& b)
{
    float r00, g00, b00, r10, g10, b10, r01, g01, b01, r11, g11, b11;
    load_rgb16f(plane, (size_t)bl.y0 * W + bl.x0, r00, g00, b00);
    load_rgb16f(plane, (size_t)bl.y0 * W + bl.x1, r10, g10, b10);
    load_rgb16f(plane, (size_t)bl.y1 * W + bl.x0, r01, g01, b01);
    load_rgb16f(plane, (size_t)bl.y1 * W + bl.x1, r11, g11, b11);
    r = bilin_mix(bl, r00, r10, r01, r11);
    g = bilin_mix(bl, g00, g10, g01, g11);
    b = bilin_mix(bl, b00, b10, b01, b11);
}

// ---- tone-map quantiser: unorm8(clamp(pow(x, .454545), 0, 1)) as a threshold search ------------------------------
// With the oracle's deterministic vk_pow the stored code is a monotone function of the single binary32 input
// (verified over ALL 2^31 - 2^23 + 1 non-negative inputs, tests/tools/pow_unorm8_sweep.c, which also generates the
// table): thr[k] is the bit pattern of the smallest x whose code is >= k.  The kernels estimate the code with the
// hardware log2 / exp2 approximations (error << 1 code) and correct it by +-1 with two exact integer compares, so
// the result is bit-identical to evaluating vk_pow and costs ~15 instead of ~60 instructions per channel.
// x must not be NaN (callers pass gl_max(0, .), which maps NaN to 0).  `thr` may point to shared or global memory.
static __device__ const uint32_t c_tonemap_thr[256] = {
#include "tonemap_thresholds.inc"
};
VK_DEVICE uint32_t tonemap_code(float x, const uint32_t* thr)
{
#ifndef VKPBRT_HOSTSIM
    const float est = exp2f(mul_rn(.454545f, __log2f(x)));      // ex2.approx(lg2.approx): pow(0) = 0, pow(inf) = inf
#else
    const float est = x > 0.0f ? exp2f(.454545f * log2f(x)) : 0.0f;
#endif
    int k = (int)fmaf(fminf(est, 1.0f), 255.0f, 0.5f);           // any rounding: only has to land within one code
    const uint32_t xb = __float_as_uint(x);
    const int kn = k < 255 ? k + 1 : 255;
    const uint32_t t_hi = thr[kn], t_lo = thr[k];
    k = (xb >= t_hi) ? kn : k;
    k = (xb < t_lo) ? k - 1 : k;                                   // thr[0] = 0: never below code 0
    return (uint32_t)k;
}
// the reference form, kept for the exhaustive device-side equivalence test (vkpbrt_debug_tonemap_sweep)
VK_DEVICE uint32_t tonemap_code_reference(float x) { return (uint32_t)f32_to_unorm8(gl_clamp(vk_pow(x, .454545f), 0.0f, 1.0f)); }

// ---- epilogue shared by bmfrPost.comp:105-123 and bfr.comp:293-308 ----
// color: clamped regression output.  Writes the rgba16f history texel and the BGRA8 tone-mapped
// texel for image pixel `pix`.
VK_DEVICE void denoise_epilogue(float cr, float cg, float cb, uint32_t frame, size_t pix, int W, int H, uint32_t motion_bits,
                                uint32_t spp_code, uchar4 alb, const uint2* __restrict__ denoised_prev,
                                uint2* __restrict__ denoised_next, uint32_t* __restrict__ final_bgra, const uint32_t* thr)
{
    float uvx = f16_bits_to_f32((uint16_t)(motion_bits & 0xffffu));
    float uvy = f16_bits_to_f32((uint16_t)(motion_bits >> 16));
    bool accept = uvx >= 0.0f;
    float pixel_spp = mul_rn(unorm8_to_f32(spp_code), 256.0f);
    float pr = 0.0f, pg = 0.0f, pb = 0.0f, blend = 1.0f;
    if (frame > 0 && accept) {
        Bilin bl = bilin_setup_repeat(uvx, uvy, W, H);
        sample_rgb16f(denoised_prev, bl, W, pr, pg, pb);
        blend = gl_max(__frcp_rn(pixel_spp), 0.1f);
    }
    float omb = sub_rn(1.0f, blend);
    cr = add_rn(mul_rn(blend, cr), mul_rn(omb, pr));
    cg = add_rn(mul_rn(blend, cg), mul_rn(omb, pg));
    cb = add_rn(mul_rn(blend, cb), mul_rn(omb, pb));
    denoised_next[pix] = pack_rgba16f(cr, cg, cb, 1.0f);
    float ar = add_rn(unorm8_to_f32(alb.x), 1e-6f), ag = add_rn(unorm8_to_f32(alb.y), 1e-6f), ab = add_rn(unorm8_to_f32(alb.z), 1e-6f);
    const uint32_t tr = tonemap_code(gl_max(0.0f, mul_rn(ar, cr)), thr);
    const uint32_t tg = tonemap_code(gl_max(0.0f, mul_rn(ag, cg)), thr);
    const uint32_t tb = tonemap_code(gl_max(0.0f, mul_rn(ab, cb)), thr);
    // B8G8R8A8_UNORM memory order
    final_bgra[pix] = tb | (tg << 8) | (tr << 16) | 0xff000000u;
}

}  // namespace vkpbrt
