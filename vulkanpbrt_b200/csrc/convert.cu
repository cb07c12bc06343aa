// convert.cu -- the small format / convention kernels on either side of the denoising path.
//
//   k_format_convert   shaders/formatConverter.comp:1-13 (host: source/renderModules/FormatConverter.cpp:4-93): the step
//                      VulkanPBRT.cpp:476-484 appends when the final image is not B8G8R8A8_UNORM (no denoiser: the path
//                      tracer's rgba32f output) -- texelFetch of the source, imageStore into the BGRA8 image.
//   k_demodulate       shaders/ptRaygen.rgen:81-88: what a producer has to emit as "demodulated illumination":
//                      min(clamp(L, 0, c_MaxRadiance) / (albedo + EPSILON), 1e3) for pixels whose primary ray hit
//                      something, the clamped radiance itself for misses (position.x is infinite there).
//   k_gbuffer_from_position_normal   source/io/RenderIO.cpp:102-120 (import) -- the conversions the reference runs on the
//                      host when it imports a BMFR-dataset style sequence: world position -> distance to the camera,
//                      cartesian normal -> (acos(n.z), atan2(n.y, n.x)), float albedo -> unorm8.
// Streaming, one pixel per thread (these are not on the per-frame hot path); non-contracted IEEE arithmetic
// (compiled with -fmad=false) so the outputs are bit-exact against oracle/vkpbrt_oracle.c.
#include "common.cuh"
#include "kernels.h"

namespace vkpbrt {

__global__ void __launch_bounds__(256) k_format_convert(const FormatConvertParams p)
{
    const int gx = blockIdx.x * 32 + threadIdx.x, gy = blockIdx.y * 8 + threadIdx.y;
    if (gx >= p.W || gy >= p.H) return;                                         // formatConverter.comp:10
    const size_t pix = (size_t)gy * p.W + gx;
    float r, g, b, a;
    if (p.src_format == 0) {                                                    // rgba32f
        const float4 v = __ldg((const float4*)p.src + pix);
        r = v.x; g = v.y; b = v.z; a = v.w;
    } else if (p.src_format == 1) {                                             // rgba16f
        const uint2 v = __ldg((const uint2*)p.src + pix);
        r = f16_bits_to_f32((uint16_t)(v.x & 0xffffu)); g = f16_bits_to_f32((uint16_t)(v.x >> 16));
        b = f16_bits_to_f32((uint16_t)(v.y & 0xffffu)); a = f16_bits_to_f32((uint16_t)(v.y >> 16));
    } else {                                                                    // rgba8 unorm (R,G,B,A bytes)
        const uint32_t v = __ldg((const uint32_t*)p.src + pix);
        r = unorm8_byte_to_f32(v, 0); g = unorm8_byte_to_f32(v, 1); b = unorm8_byte_to_f32(v, 2); a = unorm8_byte_to_f32(v, 3);
    }
    // the storage image is declared rgba8 but bound to a B8G8R8A8_UNORM view: memory order B, G, R, A
    p.dst_bgra[pix] = (uint32_t)f32_to_unorm8(b) | ((uint32_t)f32_to_unorm8(g) << 8) | ((uint32_t)f32_to_unorm8(r) << 16) |
                      ((uint32_t)f32_to_unorm8(a) << 24);
}

__global__ void __launch_bounds__(256) k_demodulate(const DemodulateParams p)
{
    const int gx = blockIdx.x * 32 + threadIdx.x, gy = blockIdx.y * 8 + threadIdx.y;
    if (gx >= p.W || gy >= p.H) return;
    const size_t pix = (size_t)gy * p.W + gx;
    const float4 L = __ldg(p.radiance + pix);
    const float4 alb = __ldg(p.albedo + pix);
    const float hit_x = __ldg(p.position_x + pix);
    float c[3] = {gl_clamp(L.x, 0.0f, 1e1f), gl_clamp(L.y, 0.0f, 1e1f), gl_clamp(L.z, 0.0f, 1e1f)};       // ptRaygen.rgen:81
    if (!isinf(hit_x)) {                                                                                  // :85
        const float al[3] = {alb.x, alb.y, alb.z};
#pragma unroll
        for (int i = 0; i < 3; ++i) c[i] = gl_min(div_rn(c[i], add_rn(al[i], 1e-6f)), 1e3f);              // :86
    }
    p.out[pix] = make_float4(c[0], c[1], c[2], 1.0f);                                                     // :88
}

// source/io/RenderIO.cpp:101-120 (world position -> distance to the eye), :160-178 (cartesian normal -> (acos(n.z),
// atan2(n.y, n.x))), :180-195 (vec4 albedo * 255.0F -> ubvec4, which TRUNCATES).  acos / atan2 are the platform's libm in
// the reference (its host code); here they are CUDA's (<= 2 ulp): the two agree to ~1e-6 rad, the tolerance of the test.
__global__ void __launch_bounds__(256) k_gbuffer_import(const GBufferImportParams p)
{
    const int gx = blockIdx.x * 32 + threadIdx.x, gy = blockIdx.y * 8 + threadIdx.y;
    if (gx >= p.W || gy >= p.H) return;
    const size_t pix = (size_t)gy * p.W + gx;
    if (p.position) {
        const float4 q = __ldg(p.position + pix);
        const float dx = sub_rn(p.camera[0], q.x), dy = sub_rn(p.camera[1], q.y), dz = sub_rn(p.camera[2], q.z);      // :116
        p.depth[pix] = sqrt_rn(add_rn(add_rn(mul_rn(dx, dx), mul_rn(dy, dy)), mul_rn(dz, dz)));
    }
    if (p.normal) {
        const float4 n = __ldg(p.normal + pix);
        p.normal_out[pix] = make_float2(acosf(n.z), atan2f(n.y, n.x));                                                // :174-175
    }
    if (p.albedo) {
        const float4 a = __ldg(p.albedo + pix);
        // float -> unsigned char conversion of an out-of-range value is undefined in C++; clamped here
        auto q8 = [](float v) { const float s = mul_rn(v, 255.0f); return (uint32_t)(s > 0.0f ? (s < 255.0f ? s : 255.0f) : 0.0f); };
        p.albedo_out[pix] = q8(a.x) | (q8(a.y) << 8) | (q8(a.z) << 16) | (q8(a.w) << 24);                                // :187
    }
}

cudaError_t launch_gbuffer_import(const GBufferImportParams& p, cudaStream_t stream)
{
    if (p.W <= 0 || p.H <= 0) return cudaSuccess;
    dim3 block(32, 8, 1), grid((p.W + 31) / 32, (p.H + 7) / 8, 1);
    VKPBRT_LAUNCH(k_gbuffer_import, grid, block, 0, stream, p);
    return cudaGetLastError();
}

cudaError_t launch_format_convert(const FormatConvertParams& p, cudaStream_t stream)
{
    if (p.W <= 0 || p.H <= 0) return cudaSuccess;
    dim3 block(32, 8, 1), grid((p.W + 31) / 32, (p.H + 7) / 8, 1);
    VKPBRT_LAUNCH(k_format_convert, grid, block, 0, stream, p);
    return cudaGetLastError();
}

cudaError_t launch_demodulate(const DemodulateParams& p, cudaStream_t stream)
{
    if (p.W <= 0 || p.H <= 0) return cudaSuccess;
    dim3 block(32, 8, 1), grid((p.W + 31) / 32, (p.H + 7) / 8, 1);
    VKPBRT_LAUNCH(k_demodulate, grid, block, 0, stream, p);
    return cudaGetLastError();
}

}  // namespace vkpbrt
