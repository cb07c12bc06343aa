// api.cpp -- host side of the C ABI declared in include/vkpbrt_b200.h.
//
// Owns the handle types (images, buffer bundles, modules), the per-frame constant folding that the
// reference leaves to the shader (accumulator.comp:46 inverse(), :54 proj * prevView) and the
// history ping-pong that replaces the reference's end-of-frame image copies.  Everything that
// touches pixels is a kernel in the .cu files next to this one; there is no CPU code path.
#include "../../include/vkpbrt_b200.h"

#ifdef VKPBRT_HOSTSIM
#include "hostsim.h"
#else
#include <cuda_runtime.h>
#endif

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <vector>

#include "kernels.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

int fail_cuda(cudaError_t e, const char* what)
{
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return VKPBRT_ERR_CUDA;
}

#define VK_CUDA(expr)                                           \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return fail_cuda(_e, #expr);     \
    } while (0)

#define VK_REQUIRE(cond, msg)                                              \
    do {                                                                   \
        if (!(cond)) return fail(VKPBRT_ERR_INVALID_ARGUMENT, (msg));      \
    } while (0)

// ---- column-major mat4 helpers, fp32, fixed evaluation order (shared definition with the oracle:
// GLSL leaves inverse()/mat*mat precision to the implementation) ------------------------------
void mat_vec(const float* m, const float* v, float* r)
{
    for (int i = 0; i < 4; ++i) r[i] = ((m[0 + i] * v[0] + m[4 + i] * v[1]) + m[8 + i] * v[2]) + m[12 + i] * v[3];
}
void mat_mul(const float* a, const float* b, float* r)
{
    for (int c = 0; c < 4; ++c) mat_vec(a, b + 4 * c, r + 4 * c);
}
void mat_inverse(const float* m, float* inv)
{
    const float a00 = m[0], a01 = m[1], a02 = m[2], a03 = m[3];
    const float a10 = m[4], a11 = m[5], a12 = m[6], a13 = m[7];
    const float a20 = m[8], a21 = m[9], a22 = m[10], a23 = m[11];
    const float a30 = m[12], a31 = m[13], a32 = m[14], a33 = m[15];
    const float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10;
    const float b03 = a01 * a12 - a02 * a11, b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
    const float b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30, b08 = a20 * a33 - a23 * a30;
    const float b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
    const float det = ((((b00 * b11 - b01 * b10) + b02 * b09) + b03 * b08) - b04 * b07) + b05 * b06;
    const float id = 1.0f / det;
    inv[0] = ((a11 * b11 - a12 * b10) + a13 * b09) * id;
    inv[1] = ((a02 * b10 - a01 * b11) - a03 * b09) * id;
    inv[2] = ((a31 * b05 - a32 * b04) + a33 * b03) * id;
    inv[3] = ((a22 * b04 - a21 * b05) - a23 * b03) * id;
    inv[4] = ((a12 * b08 - a10 * b11) - a13 * b07) * id;
    inv[5] = ((a00 * b11 - a02 * b08) + a03 * b07) * id;
    inv[6] = ((a32 * b02 - a30 * b05) - a33 * b01) * id;
    inv[7] = ((a20 * b05 - a22 * b02) + a23 * b01) * id;
    inv[8] = ((a10 * b10 - a11 * b08) + a13 * b06) * id;
    inv[9] = ((a01 * b08 - a00 * b10) - a03 * b06) * id;
    inv[10] = ((a30 * b04 - a31 * b02) + a33 * b00) * id;
    inv[11] = ((a21 * b02 - a20 * b04) - a23 * b00) * id;
    inv[12] = ((a11 * b07 - a10 * b09) - a12 * b06) * id;
    inv[13] = ((a00 * b09 - a01 * b07) + a02 * b06) * id;
    inv[14] = ((a31 * b01 - a30 * b03) - a32 * b00) * id;
    inv[15] = ((a20 * b03 - a21 * b01) + a22 * b00) * id;
}

// vsg::inverse(const mat4&), what the reference's HOST code calls (Accumulator.cpp:100 `inverse(prev.view)[3]`, the
// BMFR-dataset matrix import RenderIO.cpp:639-664): external/vsg/src/vsg/maths/maths_transform.cpp:36-156 -- affine
// matrices take t_inverse_4x3, the rest t_inverse_4x4, a zero determinant yields NaN on the diagonal; expressions in the
// source's order.  Not the shader's inverse() above.  The oracle's copy is pinned against that source (oracle/host_shim).
#define M_(c, r) m[4 * (c) + (r)]
void vsg_inverse(const float* m, float* o)
{
    const float nan = std::numeric_limits<float>::quiet_NaN();
    if (M_(0, 3) == 0.0f && M_(1, 3) == 0.0f && M_(2, 3) == 0.0f && M_(3, 3) == 1.0f) {
        const float det = (M_(0, 0) * (M_(1, 1) * M_(2, 2) - M_(1, 2) * M_(2, 1)) - M_(0, 1) * (M_(1, 0) * M_(2, 2) - M_(1, 2) * M_(2, 0))) +
                          M_(0, 2) * (M_(1, 0) * M_(2, 1) - M_(1, 1) * M_(2, 0));
        if (det == 0.0f) { for (int i = 0; i < 16; ++i) o[i] = (i % 5 == 0) ? nan : 0.0f; return; }
        const float A1223 = M_(2, 1) * M_(3, 2) - M_(2, 2) * M_(3, 1), A0223 = M_(2, 0) * M_(3, 2) - M_(2, 2) * M_(3, 0);
        const float A0123 = M_(2, 0) * M_(3, 1) - M_(2, 1) * M_(3, 0), A1213 = M_(1, 1) * M_(3, 2) - M_(1, 2) * M_(3, 1);
        const float A0213 = M_(1, 0) * M_(3, 2) - M_(1, 2) * M_(3, 0), A0113 = M_(1, 0) * M_(3, 1) - M_(1, 1) * M_(3, 0);
        const float id = 1.0f / det;
        o[0] = id * (M_(1, 1) * M_(2, 2) - M_(1, 2) * M_(2, 1));
        o[1] = id * (M_(0, 2) * M_(2, 1) - M_(0, 1) * M_(2, 2));
        o[2] = id * (M_(0, 1) * M_(1, 2) - M_(0, 2) * M_(1, 1));
        o[3] = 0.0f;
        o[4] = id * (M_(1, 2) * M_(2, 0) - M_(1, 0) * M_(2, 2));
        o[5] = id * (M_(0, 0) * M_(2, 2) - M_(0, 2) * M_(2, 0));
        o[6] = id * (M_(0, 2) * M_(1, 0) - M_(0, 0) * M_(1, 2));
        o[7] = 0.0f;
        o[8] = id * (M_(1, 0) * M_(2, 1) - M_(1, 1) * M_(2, 0));
        o[9] = id * (M_(0, 1) * M_(2, 0) - M_(0, 0) * M_(2, 1));
        o[10] = id * (M_(0, 0) * M_(1, 1) - M_(0, 1) * M_(1, 0));
        o[11] = 0.0f;
        o[12] = id * ((M_(1, 1) * A0223 - M_(1, 2) * A0123) - M_(1, 0) * A1223);
        o[13] = id * ((M_(0, 0) * A1223 - M_(0, 1) * A0223) + M_(0, 2) * A0123);
        o[14] = id * ((M_(0, 1) * A0213 - M_(0, 2) * A0113) - M_(0, 0) * A1213);
        o[15] = 1.0f;
        return;
    }
    const float A2323 = M_(2, 2) * M_(3, 3) - M_(2, 3) * M_(3, 2), A1323 = M_(2, 1) * M_(3, 3) - M_(2, 3) * M_(3, 1);
    const float A1223 = M_(2, 1) * M_(3, 2) - M_(2, 2) * M_(3, 1), A0323 = M_(2, 0) * M_(3, 3) - M_(2, 3) * M_(3, 0);
    const float A0223 = M_(2, 0) * M_(3, 2) - M_(2, 2) * M_(3, 0), A0123 = M_(2, 0) * M_(3, 1) - M_(2, 1) * M_(3, 0);
    const float A2313 = M_(1, 2) * M_(3, 3) - M_(1, 3) * M_(3, 2), A1313 = M_(1, 1) * M_(3, 3) - M_(1, 3) * M_(3, 1);
    const float A1213 = M_(1, 1) * M_(3, 2) - M_(1, 2) * M_(3, 1), A2312 = M_(1, 2) * M_(2, 3) - M_(1, 3) * M_(2, 2);
    const float A1312 = M_(1, 1) * M_(2, 3) - M_(1, 3) * M_(2, 1), A1212 = M_(1, 1) * M_(2, 2) - M_(1, 2) * M_(2, 1);
    const float A0313 = M_(1, 0) * M_(3, 3) - M_(1, 3) * M_(3, 0), A0213 = M_(1, 0) * M_(3, 2) - M_(1, 2) * M_(3, 0);
    const float A0312 = M_(1, 0) * M_(2, 3) - M_(1, 3) * M_(2, 0), A0212 = M_(1, 0) * M_(2, 2) - M_(1, 2) * M_(2, 0);
    const float A0113 = M_(1, 0) * M_(3, 1) - M_(1, 1) * M_(3, 0), A0112 = M_(1, 0) * M_(2, 1) - M_(1, 1) * M_(2, 0);
    const float det = ((M_(0, 0) * ((M_(1, 1) * A2323 - M_(1, 2) * A1323) + M_(1, 3) * A1223) - M_(0, 1) * ((M_(1, 0) * A2323 - M_(1, 2) * A0323) + M_(1, 3) * A0223)) +
                       M_(0, 2) * ((M_(1, 0) * A1323 - M_(1, 1) * A0323) + M_(1, 3) * A0123)) -
                      M_(0, 3) * ((M_(1, 0) * A1223 - M_(1, 1) * A0223) + M_(1, 2) * A0123);
    if (det == 0.0f) { for (int i = 0; i < 16; ++i) o[i] = (i % 5 == 0) ? nan : 0.0f; return; }
    const float id = 1.0f / det;
    o[0] = id * ((M_(1, 1) * A2323 - M_(1, 2) * A1323) + M_(1, 3) * A1223);
    o[1] = id * -((M_(0, 1) * A2323 - M_(0, 2) * A1323) + M_(0, 3) * A1223);
    o[2] = id * ((M_(0, 1) * A2313 - M_(0, 2) * A1313) + M_(0, 3) * A1213);
    o[3] = id * -((M_(0, 1) * A2312 - M_(0, 2) * A1312) + M_(0, 3) * A1212);
    o[4] = id * -((M_(1, 0) * A2323 - M_(1, 2) * A0323) + M_(1, 3) * A0223);
    o[5] = id * ((M_(0, 0) * A2323 - M_(0, 2) * A0323) + M_(0, 3) * A0223);
    o[6] = id * -((M_(0, 0) * A2313 - M_(0, 2) * A0313) + M_(0, 3) * A0213);
    o[7] = id * ((M_(0, 0) * A2312 - M_(0, 2) * A0312) + M_(0, 3) * A0212);
    o[8] = id * ((M_(1, 0) * A1323 - M_(1, 1) * A0323) + M_(1, 3) * A0123);
    o[9] = id * -((M_(0, 0) * A1323 - M_(0, 1) * A0323) + M_(0, 3) * A0123);
    o[10] = id * ((M_(0, 0) * A1313 - M_(0, 1) * A0313) + M_(0, 3) * A0113);
    o[11] = id * -((M_(0, 0) * A1312 - M_(0, 1) * A0312) + M_(0, 3) * A0112);
    o[12] = id * -((M_(1, 0) * A1223 - M_(1, 1) * A0223) + M_(1, 2) * A0123);
    o[13] = id * ((M_(0, 0) * A1223 - M_(0, 1) * A0223) + M_(0, 2) * A0123);
    o[14] = id * -((M_(0, 0) * A1213 - M_(0, 1) * A0213) + M_(0, 2) * A0113);
    o[15] = id * ((M_(0, 0) * A1212 - M_(0, 1) * A0212) + M_(0, 2) * A0112);
}
#undef M_

}  // namespace

// ------------------------------------------------------------------------------------------------
// handle types
// ------------------------------------------------------------------------------------------------
struct vkpbrt_context_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::atomic<uint64_t> launches{0};
    // side lanes: independent denoisers of one frame (the three block sizes of X8X16X32) run concurrently on their
    // own streams, forked from / joined back into `stream` with events
    static constexpr int kLanes = 2;
    cudaStream_t lane[kLanes] = {nullptr, nullptr};
    cudaEvent_t fork_ev[kLanes] = {nullptr, nullptr}, join_ev[kLanes] = {nullptr, nullptr};
    bool lane_pending[kLanes] = {false, false};
};

struct vkpbrt_image_s {
    vkpbrt_context_t ctx = nullptr;
    uint32_t format = 0, width = 0, height = 0, layers = 1;
    void* data = nullptr;
    void* allocation = nullptr;   // non-null when owned
    bool owned = true;
    std::atomic<int> refs{1};
    uint64_t row_pitch() const { return (uint64_t)width * vkpbrt_format_texel_size(format); }
    uint64_t layer_pitch() const { return row_pitch() * height; }
    uint64_t size_bytes() const { return layer_pitch() * layers; }
};

struct vkpbrt_gbuffer_s {
    vkpbrt_context_t ctx;
    uint32_t width, height;
    vkpbrt_image_t img[4];
};

struct vkpbrt_illumination_buffer_s {
    vkpbrt_context_t ctx;
    uint32_t type, width, height;
    std::vector<vkpbrt_image_t> images;
};

struct vkpbrt_accumulation_buffer_s {
    vkpbrt_context_t ctx;
    uint32_t width, height;
    vkpbrt_image_t img[7];
    vkpbrt_image_t depth_next = nullptr;   // ping-pong partner of prev_depth, written by k_accumulate
    bool depth_next_valid = false;
};

struct vkpbrt_accumulator_s {
    vkpbrt_context_t ctx;
    int width, height, work_width, work_height;
    bool separate_matrices;
    vkpbrt_gbuffer_t g;
    vkpbrt_illumination_buffer_t original;
    vkpbrt_illumination_buffer_t accumulated;
    vkpbrt_accumulation_buffer_t acc;
    bool compiled = false;
    int row_begin, row_end;
    // Accumulator::PushConstants (Accumulator.hpp:28-33)
    float pc_view[16], pc_inv_view[16], pc_prev_view[16], pc_prev_pos[4];
    int pc_frame_number = 0;
    bool force_scalar = false;
    int max_disp_rows = 0;
    uint32_t* disp_violations = nullptr;    // device word, allocated when the guard is switched on
};

struct vkpbrt_bmfr_s {
    vkpbrt_context_t ctx;
    uint32_t width, height, work, fitting_kernel, blocks_x, blocks_y;
    vkpbrt_gbuffer_t g;
    vkpbrt_illumination_buffer_t illum;
    vkpbrt_accumulation_buffer_t acc;
    vkpbrt_image_t denoised = nullptr, final_image = nullptr, features = nullptr, weights = nullptr;
    bool debug = false, force_generic = false, compiled = false;
    int block_row_begin, block_row_end;
    // block-invariant per-frame table (csrc/bmfr.cu), double-buffered: slot f & 1 holds frame table_frame[f & 1]
    float* table[2] = {nullptr, nullptr};
    int64_t table_frame[2] = {-1, -1};
    int lane = 0;                 // as vkpbrt_bfr_s::lane
    int position_type = 0;        // bmfrGeneral.comp:30-31 POSITION_TYPE
    bool tma_enabled = true;      // interior blocks stage their input tiles through TMA (VKPBRT_BMFR_TMA=0 switches it off)
};

struct vkpbrt_bfr_s {
    vkpbrt_context_t ctx;
    uint32_t width, height, work, blocks_x, blocks_y;
    vkpbrt_gbuffer_t g;
    vkpbrt_illumination_buffer_t illum;
    vkpbrt_accumulation_buffer_t acc;
    vkpbrt_image_t denoised = nullptr, final_image = nullptr;
    bool compiled = false;
    int lane = 0;                 // 0: the context's stream; 1, 2: a side lane (runs concurrently with the other denoisers of the frame)
    float lr_exp[40], lr_sqrt[40], lr_den[40];
};

struct vkpbrt_bfr_blender_s {
    vkpbrt_context_t ctx;
    uint32_t width, height, work_width, work_height, radius;
    vkpbrt_image_t average, average_squared, den[3];
    vkpbrt_image_t final_image = nullptr;
    bool compiled = false;
};

struct vkpbrt_taa_s {
    vkpbrt_context_t ctx;
    uint32_t width, height, work_width, work_height;
    vkpbrt_accumulation_buffer_t acc;
    vkpbrt_image_t denoised;
    vkpbrt_image_t final_image = nullptr, history = nullptr;   // handles; data flips between buf[0..1]
    void* buf[2] = {nullptr, nullptr};
    int strip_rows = 0;           // test switch: strip height of the pair kernel (0 = chosen per launch)
    int fix_swizzle = 0;
    int force_scalar = 0;
    bool compiled = false;
    int row_begin, row_end;
};

struct vkpbrt_format_converter_s {
    vkpbrt_context_t ctx;
    vkpbrt_image_t src;
    vkpbrt_image_t final_image = nullptr;
    uint32_t work_width, work_height;
    bool compiled = false;
};

struct vkpbrt_external_memory_s {
    vkpbrt_context_t ctx;
    cudaExternalMemory_t mem;
    void* ptr;
};
struct vkpbrt_external_semaphore_s {
    vkpbrt_context_t ctx;
    cudaExternalSemaphore_t sem;
    bool timeline;
};

// the context's stream with every side lane joined back into it: all work recorded so far is ordered before whatever
// the caller enqueues next
static cudaStream_t joined(vkpbrt_context_t ctx)
{
    for (int i = 0; i < vkpbrt_context_s::kLanes; ++i)
        if (ctx->lane_pending[i]) {
            cudaStreamWaitEvent(ctx->stream, ctx->join_ev[i], 0);
            ctx->lane_pending[i] = false;
        }
    return ctx->stream;
}
// stream of lane `lane` (1-based; 0 = the context's stream), ordered after everything recorded on the context's stream
static cudaError_t fork_lane(vkpbrt_context_t ctx, int lane, cudaStream_t* out)
{
    if (lane <= 0 || lane > vkpbrt_context_s::kLanes) { *out = joined(ctx); return cudaSuccess; }
    const int i = lane - 1;
    cudaError_t e = cudaSuccess;
    if (!ctx->lane[i]) {
        e = cudaStreamCreateWithFlags(&ctx->lane[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->fork_ev[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->join_ev[i], cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    // the lane's previous work (last frame) is already joined or still pending: either way stream order on the lane
    // itself keeps its launches in sequence
    e = cudaEventRecord(ctx->fork_ev[i], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->lane[i], ctx->fork_ev[i], 0);
    *out = ctx->lane[i];
    return e;
}
static cudaError_t lane_done(vkpbrt_context_t ctx, int lane)
{
    if (lane <= 0 || lane > vkpbrt_context_s::kLanes) return cudaSuccess;
    const int i = lane - 1;
    cudaError_t e = cudaEventRecord(ctx->join_ev[i], ctx->lane[i]);
    if (e == cudaSuccess) ctx->lane_pending[i] = true;
    return e;
}

// ------------------------------------------------------------------------------------------------
extern "C" {

const char* vkpbrt_last_error(void) { return g_last_error.c_str(); }
const char* vkpbrt_version(void)
{
#ifdef VKPBRT_HOSTSIM
    return "vkpbrt_b200 HOSTSIM (CPU test emulator, not a product build)";
#else
    return "vkpbrt_b200 0.1 (sm_100a)";
#endif
}

uint32_t vkpbrt_format_texel_size(uint32_t format)
{
    switch (format) {
    case VKPBRT_FORMAT_R32_SFLOAT: return 4;
    case VKPBRT_FORMAT_R32G32_SFLOAT: return 8;
    case VKPBRT_FORMAT_R8G8B8A8_UNORM: return 4;
    case VKPBRT_FORMAT_B8G8R8A8_UNORM: return 4;
    case VKPBRT_FORMAT_R16G16_SFLOAT: return 4;
    case VKPBRT_FORMAT_R8_UNORM: return 1;
    case VKPBRT_FORMAT_R16G16B16A16_SFLOAT: return 8;
    case VKPBRT_FORMAT_R32G32B32A32_SFLOAT: return 16;
    case VKPBRT_FORMAT_R16_SFLOAT: return 2;
    default: return 0;
    }
}

// ---- context -----------------------------------------------------------------------------------
int vkpbrt_context_create(int device, void* cuda_stream, vkpbrt_context_t* out)
{
    VK_REQUIRE(out, "vkpbrt_context_create: out is null");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(VKPBRT_ERR_NO_DEVICE, std::string("vkpbrt_context_create: no CUDA device (") + cudaGetErrorString(e) +
                                              "); this library has no CPU fallback");
    VK_REQUIRE(device >= 0 && device < count, "vkpbrt_context_create: device index out of range");
    cudaDeviceProp prop;
    VK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(VKPBRT_ERR_NO_DEVICE, std::string("vkpbrt_context_create: device '") + prop.name + "' is sm_" +
                                              std::to_string(prop.major) + std::to_string(prop.minor) +
                                              "; kernels are built for sm_100a only");
    VK_CUDA(cudaSetDevice(device));
    auto* c = new (std::nothrow) vkpbrt_context_s();
    VK_REQUIRE(c, "out of host memory");
    c->device = device;
    if (cuda_stream) {
        c->stream = (cudaStream_t)cuda_stream;
    } else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail_cuda(e, "cudaStreamCreateWithFlags"); }
        c->own_stream = true;
    }
    *out = c;
    return VKPBRT_OK;
}

int vkpbrt_context_destroy(vkpbrt_context_t ctx)
{
    if (!ctx) return VKPBRT_OK;
    for (int i = 0; i < vkpbrt_context_s::kLanes; ++i) {
        if (ctx->lane[i]) cudaStreamDestroy(ctx->lane[i]);
        if (ctx->fork_ev[i]) cudaEventDestroy(ctx->fork_ev[i]);
        if (ctx->join_ev[i]) cudaEventDestroy(ctx->join_ev[i]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return VKPBRT_OK;
}

int vkpbrt_context_synchronize(vkpbrt_context_t ctx)
{
    VK_REQUIRE(ctx, "null context");
    VK_CUDA(cudaStreamSynchronize(joined(ctx)));
    return VKPBRT_OK;
}

int vkpbrt_context_stream(vkpbrt_context_t ctx, void** cuda_stream)
{
    VK_REQUIRE(ctx && cuda_stream, "null argument");
    *cuda_stream = (void*)ctx->stream;
    return VKPBRT_OK;
}

int vkpbrt_context_launch_count(vkpbrt_context_t ctx, uint64_t* out)
{
    VK_REQUIRE(ctx && out, "null argument");
    *out = ctx->launches.load();
    return VKPBRT_OK;
}

int vkpbrt_mat4_inverse(const float* m, float* inverse)
{
    VK_REQUIRE(m && inverse, "vkpbrt_mat4_inverse: null argument");
    float tmp[16];
    vsg_inverse(m, tmp);
    std::memcpy(inverse, tmp, sizeof tmp);
    return VKPBRT_OK;
}

int vkpbrt_device_count(int* count)
{
    VK_REQUIRE(count, "vkpbrt_device_count: null argument");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { *count = 0; return fail(VKPBRT_ERR_NO_DEVICE, std::string("vkpbrt_device_count: ") + cudaGetErrorString(e)); }
    return VKPBRT_OK;
}

int vkpbrt_device_uuid(int device, uint8_t uuid[16])
{
    VK_REQUIRE(uuid, "vkpbrt_device_uuid: null argument");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(VKPBRT_ERR_NO_DEVICE, std::string("vkpbrt_device_uuid: no CUDA device (") + cudaGetErrorString(e) + ")");
    VK_REQUIRE(device >= 0 && device < count, "vkpbrt_device_uuid: device index out of range");
    cudaDeviceProp prop;
    VK_CUDA(cudaGetDeviceProperties(&prop, device));
    static_assert(sizeof(prop.uuid.bytes) == 16, "CUDA device UUIDs are 16 bytes, like VkPhysicalDeviceIDProperties::deviceUUID");
    memcpy(uuid, prop.uuid.bytes, 16);
    return VKPBRT_OK;
}

// ---- device-side self checks ---------------------------------------------------------------------
int vkpbrt_debug_tonemap_sweep(vkpbrt_context_t ctx, uint64_t* mismatches, uint32_t* first_mismatch)
{
    VK_REQUIRE(ctx && mismatches && first_mismatch, "vkpbrt_debug_tonemap_sweep: null argument");
    VK_CUDA(cudaSetDevice(ctx->device));
    unsigned long long* d_bad = nullptr;
    VK_CUDA(cudaMalloc((void**)&d_bad, 16));
    uint32_t* d_first = reinterpret_cast<uint32_t*>(d_bad + 1);
    const unsigned long long zero = 0;
    const uint32_t none = 0xffffffffu;
    VK_CUDA(cudaMemcpyAsync(d_bad, &zero, 8, cudaMemcpyHostToDevice, joined(ctx)));
    VK_CUDA(cudaMemcpyAsync(d_first, &none, 4, cudaMemcpyHostToDevice, joined(ctx)));
    cudaError_t e = vkpbrt::launch_tonemap_sweep(d_bad, d_first, joined(ctx));
    unsigned long long bad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, joined(ctx));
    if (e == cudaSuccess) e = cudaMemcpyAsync(first_mismatch, d_first, 4, cudaMemcpyDeviceToHost, joined(ctx));
    if (e == cudaSuccess) e = cudaStreamSynchronize(joined(ctx));
    cudaFree(d_bad);
    VK_CUDA(e);
    *mismatches = bad;
    ctx->launches++;
    return VKPBRT_OK;
}

// ---- images ------------------------------------------------------------------------------------
int vkpbrt_image_create(vkpbrt_context_t ctx, uint32_t format, uint32_t width, uint32_t height, uint32_t layers,
                        vkpbrt_image_t* out)
{
    VK_REQUIRE(ctx && out, "vkpbrt_image_create: null argument");
    VK_REQUIRE(vkpbrt_format_texel_size(format) != 0, "vkpbrt_image_create: unknown format");
    VK_REQUIRE(width > 0 && height > 0 && layers > 0, "vkpbrt_image_create: empty extent");
    auto* i = new (std::nothrow) vkpbrt_image_s();
    VK_REQUIRE(i, "out of host memory");
    i->ctx = ctx; i->format = format; i->width = width; i->height = height; i->layers = layers;
    i->owned = true;
    *out = i;
    return VKPBRT_OK;
}

int vkpbrt_image_wrap(vkpbrt_context_t ctx, uint32_t format, uint32_t width, uint32_t height, uint32_t layers,
                      void* device_ptr, vkpbrt_image_t* out)
{
    int rc = vkpbrt_image_create(ctx, format, width, height, layers, out);
    if (rc) return rc;
    (*out)->owned = false;
    (*out)->data = device_ptr;
    return VKPBRT_OK;
}

int vkpbrt_image_set_data(vkpbrt_image_t img, void* device_ptr)
{
    VK_REQUIRE(img, "null image");
    VK_REQUIRE(!img->owned, "vkpbrt_image_set_data: image owns its memory");
    img->data = device_ptr;
    return VKPBRT_OK;
}

int vkpbrt_image_compile(vkpbrt_image_t img)
{
    VK_REQUIRE(img, "null image");
    if (!img->owned) {
        VK_REQUIRE(img->data, "vkpbrt_image_compile: wrapped image has no memory");
        return VKPBRT_OK;
    }
    if (img->allocation) return VKPBRT_OK;
    VK_CUDA(cudaSetDevice(img->ctx->device));
    VK_CUDA(cudaMalloc(&img->allocation, img->size_bytes()));
    img->data = img->allocation;
    // the reference leaves new images undefined; zero is the documented initial history
    VK_CUDA(cudaMemsetAsync(img->data, 0, img->size_bytes(), joined(img->ctx)));
    return VKPBRT_OK;
}

int vkpbrt_image_info_get(vkpbrt_image_t img, vkpbrt_image_info* out)
{
    VK_REQUIRE(img && out, "null argument");
    out->data = img->data;
    out->format = img->format;
    out->width = img->width; out->height = img->height; out->layers = img->layers;
    out->row_pitch = img->row_pitch();
    out->layer_pitch = img->layer_pitch();
    out->size_bytes = img->size_bytes();
    out->owned = img->owned ? 1 : 0;
    return VKPBRT_OK;
}

int vkpbrt_image_upload(vkpbrt_image_t img, const void* host, uint64_t bytes)
{
    VK_REQUIRE(img && host, "null argument");
    VK_REQUIRE(img->data, "vkpbrt_image_upload: image not compiled");
    VK_REQUIRE(bytes <= img->size_bytes(), "vkpbrt_image_upload: size exceeds image");
    VK_CUDA(cudaMemcpyAsync(img->data, host, bytes, cudaMemcpyHostToDevice, joined(img->ctx)));
    return VKPBRT_OK;
}

int vkpbrt_image_download(vkpbrt_image_t img, void* host, uint64_t bytes)
{
    VK_REQUIRE(img && host, "null argument");
    VK_REQUIRE(img->data, "vkpbrt_image_download: image not compiled");
    VK_REQUIRE(bytes <= img->size_bytes(), "vkpbrt_image_download: size exceeds image");
    VK_CUDA(cudaMemcpyAsync(host, img->data, bytes, cudaMemcpyDeviceToHost, joined(img->ctx)));
    return VKPBRT_OK;
}

int vkpbrt_image_clear(vkpbrt_image_t img)
{
    VK_REQUIRE(img && img->data, "vkpbrt_image_clear: image not compiled");
    VK_CUDA(cudaMemsetAsync(img->data, 0, img->size_bytes(), joined(img->ctx)));
    return VKPBRT_OK;
}

int vkpbrt_image_copy_record(vkpbrt_image_t src, vkpbrt_image_t dst)
{
    VK_REQUIRE(src && dst, "vkpbrt_image_copy_record: null image");
    VK_REQUIRE(src->data && dst->data, "vkpbrt_image_copy_record: image not compiled");
    VK_REQUIRE(src->ctx == dst->ctx, "vkpbrt_image_copy_record: images of different contexts");
    VK_REQUIRE(src->width == dst->width && src->height == dst->height && src->layers == dst->layers,
               "vkpbrt_image_copy_record: extents differ");
    VK_REQUIRE(vkpbrt_format_texel_size(src->format) == vkpbrt_format_texel_size(dst->format),
               "vkpbrt_image_copy_record: texel sizes differ (vkCmdCopyImage requires size-compatible formats)");
    if (src->data == dst->data) return VKPBRT_OK;
    VK_CUDA(cudaSetDevice(src->ctx->device));
    VK_CUDA(cudaMemcpyAsync(dst->data, src->data, src->size_bytes(), cudaMemcpyDeviceToDevice, joined(src->ctx)));
    return VKPBRT_OK;
}

int vkpbrt_image_retain(vkpbrt_image_t img)
{
    VK_REQUIRE(img, "null image");
    img->refs.fetch_add(1);
    return VKPBRT_OK;
}

int vkpbrt_image_release(vkpbrt_image_t img)
{
    if (!img) return VKPBRT_OK;
    if (img->refs.fetch_sub(1) == 1) {
        if (img->allocation) cudaFree(img->allocation);
        delete img;
    }
    return VKPBRT_OK;
}

static int make_image(vkpbrt_context_t ctx, uint32_t fmt, uint32_t w, uint32_t h, uint32_t layers, vkpbrt_image_t* out)
{
    return vkpbrt_image_create(ctx, fmt, w, h, layers, out);
}

static bool same_extent(vkpbrt_image_t i, uint32_t w, uint32_t h) { return i && i->width == w && i->height == h; }

// ---- GBuffer (source/buffers/GBuffer.cpp:53-125) -------------------------------------------------
int vkpbrt_gbuffer_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, vkpbrt_gbuffer_t* out)
{
    VK_REQUIRE(ctx && out, "null argument");
    auto* g = new vkpbrt_gbuffer_s{ctx, width, height, {nullptr, nullptr, nullptr, nullptr}};
    const uint32_t fmts[4] = {VKPBRT_FORMAT_R32_SFLOAT, VKPBRT_FORMAT_R32G32_SFLOAT, VKPBRT_FORMAT_R8G8B8A8_UNORM,
                              VKPBRT_FORMAT_R8G8B8A8_UNORM};
    for (int i = 0; i < 4; ++i) {
        int rc = make_image(ctx, fmts[i], width, height, 1, &g->img[i]);
        if (rc) { vkpbrt_gbuffer_destroy(g); return rc; }
    }
    *out = g;
    return VKPBRT_OK;
}

int vkpbrt_gbuffer_create_from_images(vkpbrt_context_t ctx, vkpbrt_image_t depth, vkpbrt_image_t normal,
                                      vkpbrt_image_t material, vkpbrt_image_t albedo, vkpbrt_gbuffer_t* out)
{
    VK_REQUIRE(ctx && out && depth && normal && albedo, "vkpbrt_gbuffer_create_from_images: null argument");
    VK_REQUIRE(depth->format == VKPBRT_FORMAT_R32_SFLOAT, "depth must be R32_SFLOAT (GBuffer.cpp:60)");
    VK_REQUIRE(normal->format == VKPBRT_FORMAT_R32G32_SFLOAT, "normal must be R32G32_SFLOAT (GBuffer.cpp:77)");
    VK_REQUIRE(albedo->format == VKPBRT_FORMAT_R8G8B8A8_UNORM, "albedo must be R8G8B8A8_UNORM (GBuffer.cpp:111)");
    VK_REQUIRE(same_extent(normal, depth->width, depth->height) && same_extent(albedo, depth->width, depth->height),
               "g-buffer planes differ in extent");
    auto* g = new vkpbrt_gbuffer_s{ctx, depth->width, depth->height, {depth, normal, material, albedo}};
    for (auto* i : g->img)
        if (i) vkpbrt_image_retain(i);
    *out = g;
    return VKPBRT_OK;
}

int vkpbrt_gbuffer_compile(vkpbrt_gbuffer_t g)
{
    VK_REQUIRE(g, "null gbuffer");
    for (auto* i : g->img)
        if (i) { int rc = vkpbrt_image_compile(i); if (rc) return rc; }
    return VKPBRT_OK;
}

int vkpbrt_gbuffer_image(vkpbrt_gbuffer_t g, uint32_t member, vkpbrt_image_t* out)
{
    VK_REQUIRE(g && out && member < 4, "vkpbrt_gbuffer_image: bad argument");
    *out = g->img[member];
    return VKPBRT_OK;
}

int vkpbrt_gbuffer_destroy(vkpbrt_gbuffer_t g)
{
    if (!g) return VKPBRT_OK;
    for (auto* i : g->img) vkpbrt_image_release(i);
    delete g;
    return VKPBRT_OK;
}

// ---- IlluminationBuffer (source/buffers/IlluminationBuffer.cpp:223-282) --------------------------
int vkpbrt_illumination_buffer_create(vkpbrt_context_t ctx, uint32_t type, uint32_t width, uint32_t height,
                                      vkpbrt_illumination_buffer_t* out)
{
    VK_REQUIRE(ctx && out, "null argument");
    auto* b = new vkpbrt_illumination_buffer_s{ctx, type, width, height, {}};
    std::vector<uint32_t> fmts;
    switch (type) {
    case VKPBRT_ILLUMINATION_DEMODULATED: fmts = {VKPBRT_FORMAT_R16G16B16A16_SFLOAT, VKPBRT_FORMAT_R16G16B16A16_SFLOAT}; break;
    case VKPBRT_ILLUMINATION_DEMODULATED_FLOAT: fmts = {VKPBRT_FORMAT_R32G32B32A32_SFLOAT}; break;
    case VKPBRT_ILLUMINATION_FINAL: fmts = {VKPBRT_FORMAT_B8G8R8A8_UNORM}; break;              // IlluminationBuffer.cpp:84
    case VKPBRT_ILLUMINATION_FINAL_DEMODULATED:                                                   // IlluminationBuffer.cpp:174
        fmts = {VKPBRT_FORMAT_B8G8R8A8_UNORM, VKPBRT_FORMAT_R16G16B16A16_SFLOAT, VKPBRT_FORMAT_R16G16B16A16_SFLOAT};
        break;
    default: delete b; return fail(VKPBRT_ERR_INVALID_ARGUMENT, "unknown illumination buffer type");
    }
    for (uint32_t f : fmts) {
        vkpbrt_image_t i = nullptr;
        int rc = make_image(ctx, f, width, height, 1, &i);
        if (rc) { vkpbrt_illumination_buffer_destroy(b); return rc; }
        b->images.push_back(i);
    }
    *out = b;
    return VKPBRT_OK;
}

int vkpbrt_illumination_buffer_create_from_images(vkpbrt_context_t ctx, uint32_t type, const vkpbrt_image_t* images,
                                                  uint32_t count, vkpbrt_illumination_buffer_t* out)
{
    VK_REQUIRE(ctx && images && out && count > 0, "vkpbrt_illumination_buffer_create_from_images: bad argument");
    VK_REQUIRE(type <= VKPBRT_ILLUMINATION_FINAL_DEMODULATED, "unknown illumination buffer type");
    for (uint32_t i = 0; i < count; ++i) VK_REQUIRE(images[i], "null image");
    auto* b = new vkpbrt_illumination_buffer_s{ctx, type, images[0]->width, images[0]->height, {}};
    for (uint32_t i = 0; i < count; ++i) {
        vkpbrt_image_retain(images[i]);
        b->images.push_back(images[i]);
    }
    *out = b;
    return VKPBRT_OK;
}

int vkpbrt_illumination_buffer_compile(vkpbrt_illumination_buffer_t b)
{
    VK_REQUIRE(b, "null illumination buffer");
    for (auto* i : b->images) { int rc = vkpbrt_image_compile(i); if (rc) return rc; }
    return VKPBRT_OK;
}

int vkpbrt_illumination_buffer_type(vkpbrt_illumination_buffer_t b, uint32_t* type, uint32_t* image_count)
{
    VK_REQUIRE(b, "null illumination buffer");
    if (type) *type = b->type;
    if (image_count) *image_count = (uint32_t)b->images.size();
    return VKPBRT_OK;
}

int vkpbrt_illumination_buffer_image(vkpbrt_illumination_buffer_t b, uint32_t index, vkpbrt_image_t* out)
{
    VK_REQUIRE(b && out, "null argument");
    VK_REQUIRE(index < b->images.size(), "illumination_images index out of range");
    *out = b->images[index];
    return VKPBRT_OK;
}

int vkpbrt_illumination_buffer_destroy(vkpbrt_illumination_buffer_t b)
{
    if (!b) return VKPBRT_OK;
    for (auto* i : b->images) vkpbrt_image_release(i);
    delete b;
    return VKPBRT_OK;
}

// ---- AccumulationBuffer (source/buffers/AccumulationBuffer.cpp:245-339) --------------------------
int vkpbrt_accumulation_buffer_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, vkpbrt_accumulation_buffer_t* out)
{
    VK_REQUIRE(ctx && out, "null argument");
    auto* b = new vkpbrt_accumulation_buffer_s();
    b->ctx = ctx; b->width = width; b->height = height;
    for (auto& i : b->img) i = nullptr;
    const uint32_t fmts[7] = {VKPBRT_FORMAT_R16G16B16A16_SFLOAT, VKPBRT_FORMAT_R16G16B16A16_SFLOAT, VKPBRT_FORMAT_R32_SFLOAT,
                              VKPBRT_FORMAT_R32G32_SFLOAT, VKPBRT_FORMAT_R8_UNORM, VKPBRT_FORMAT_R8_UNORM,
                              VKPBRT_FORMAT_R16G16_SFLOAT};
    for (int i = 0; i < 7; ++i) {
        int rc = make_image(ctx, fmts[i], width, height, 1, &b->img[i]);
        if (rc) { vkpbrt_accumulation_buffer_destroy(b); return rc; }
    }
    int rc = make_image(ctx, VKPBRT_FORMAT_R32_SFLOAT, width, height, 1, &b->depth_next);
    if (rc) { vkpbrt_accumulation_buffer_destroy(b); return rc; }
    *out = b;
    return VKPBRT_OK;
}

int vkpbrt_accumulation_buffer_compile(vkpbrt_accumulation_buffer_t b)
{
    VK_REQUIRE(b, "null accumulation buffer");
    for (int i = 0; i < 7; ++i) {
        // prev_normal and prev_illu_squared are never read on the path: keep the handles, skip the memory
        if (i == VKPBRT_ACC_PREV_NORMAL || i == VKPBRT_ACC_PREV_ILLU_SQUARED) continue;
        int rc = vkpbrt_image_compile(b->img[i]);
        if (rc) return rc;
    }
    return vkpbrt_image_compile(b->depth_next);
}

int vkpbrt_accumulation_buffer_image(vkpbrt_accumulation_buffer_t b, uint32_t member, vkpbrt_image_t* out)
{
    VK_REQUIRE(b && out && member < 8, "vkpbrt_accumulation_buffer_image: bad argument");
    *out = member == VKPBRT_ACC_NEXT_DEPTH ? b->depth_next : b->img[member];
    return VKPBRT_OK;
}

static int swap_or_copy(vkpbrt_image_t cur, vkpbrt_image_t prev)
{
    VK_REQUIRE(cur->data && prev->data, "copy_to_back_images: image not compiled");
    VK_REQUIRE(cur->size_bytes() == prev->size_bytes(), "copy_to_back_images: extent mismatch");
    if (cur->owned && prev->owned) {
        std::swap(cur->data, prev->data);
        std::swap(cur->allocation, prev->allocation);
    } else {
        VK_CUDA(cudaMemcpyAsync(prev->data, cur->data, cur->size_bytes(), cudaMemcpyDeviceToDevice, joined(cur->ctx)));
    }
    return VKPBRT_OK;
}

int vkpbrt_accumulation_buffer_copy_to_back_images(vkpbrt_accumulation_buffer_t b, vkpbrt_gbuffer_t g,
                                                   vkpbrt_illumination_buffer_t illumination)
{
    VK_REQUIRE(b && g && illumination, "copy_to_back_images: null argument");
    if (illumination->type != VKPBRT_ILLUMINATION_DEMODULATED)   // AccumulationBuffer.cpp:178, :217 throw
        return fail(VKPBRT_ERR_UNSUPPORTED, "AccumulationBuffer::copy_to_back_images: illumination buffer not supported");
    // depth -> prev_depth
    if (b->depth_next_valid) {
        std::swap(b->img[VKPBRT_ACC_PREV_DEPTH]->data, b->depth_next->data);
        std::swap(b->img[VKPBRT_ACC_PREV_DEPTH]->allocation, b->depth_next->allocation);
        b->depth_next_valid = false;
    } else {
        vkpbrt_image_t d = g->img[VKPBRT_GBUFFER_DEPTH], pd = b->img[VKPBRT_ACC_PREV_DEPTH];
        VK_REQUIRE(d->data && pd->data, "copy_to_back_images: depth image not compiled");
        VK_CUDA(cudaMemcpyAsync(pd->data, d->data, pd->size_bytes(), cudaMemcpyDeviceToDevice, joined(b->ctx)));
    }
    int rc = swap_or_copy(b->img[VKPBRT_ACC_SPP], b->img[VKPBRT_ACC_PREV_SPP]);
    if (rc) return rc;
    return swap_or_copy(illumination->images[0], b->img[VKPBRT_ACC_PREV_ILLU]);
}

int vkpbrt_accumulation_buffer_destroy(vkpbrt_accumulation_buffer_t b)
{
    if (!b) return VKPBRT_OK;
    for (auto* i : b->img) vkpbrt_image_release(i);
    vkpbrt_image_release(b->depth_next);
    delete b;
    return VKPBRT_OK;
}

// ---- Accumulator ---------------------------------------------------------------------------------
int vkpbrt_accumulator_create(vkpbrt_context_t ctx, vkpbrt_gbuffer_t g, vkpbrt_illumination_buffer_t illumination,
                              int separate_matrices, int work_width, int work_height, vkpbrt_accumulator_t* out)
{
    VK_REQUIRE(ctx && g && illumination && out, "vkpbrt_accumulator_create: null argument");
    VK_REQUIRE(!illumination->images.empty(), "illumination buffer has no images");
    const uint32_t sf = illumination->images[0]->format;
    if (sf != VKPBRT_FORMAT_R32G32B32A32_SFLOAT && sf != VKPBRT_FORMAT_R16G16B16A16_SFLOAT)
        return fail(VKPBRT_ERR_UNSUPPORTED, "Accumulator: illumination_images[0] must be rgba32f or rgba16f");
    VK_REQUIRE(same_extent(illumination->images[0], g->width, g->height), "illumination / g-buffer extent mismatch");
    auto* a = new vkpbrt_accumulator_s();
    a->ctx = ctx;
    a->width = (int)g->width; a->height = (int)g->height;       // Accumulator.cpp:6-7
    a->work_width = work_width; a->work_height = work_height;
    a->separate_matrices = separate_matrices != 0;
    a->g = g; a->original = illumination;
    a->accumulated = nullptr; a->acc = nullptr;
    a->row_begin = 0; a->row_end = a->height;
    std::memset(a->pc_view, 0, sizeof a->pc_view);
    std::memset(a->pc_inv_view, 0, sizeof a->pc_inv_view);
    std::memset(a->pc_prev_view, 0, sizeof a->pc_prev_view);
    std::memset(a->pc_prev_pos, 0, sizeof a->pc_prev_pos);
    int rc = vkpbrt_illumination_buffer_create(ctx, VKPBRT_ILLUMINATION_DEMODULATED, g->width, g->height, &a->accumulated);   // :8
    if (!rc) rc = vkpbrt_accumulation_buffer_create(ctx, g->width, g->height, &a->acc);                                       // :9
    if (rc) { vkpbrt_accumulator_destroy(a); return rc; }
    *out = a;
    return VKPBRT_OK;
}

int vkpbrt_accumulator_compile_images(vkpbrt_accumulator_t a)
{
    VK_REQUIRE(a, "null accumulator");
    int rc = vkpbrt_accumulation_buffer_compile(a->acc);
    // illuminationSquared (images[1]) is declared but never written by any shader (accumulator.comp:104
    // is the only illumination store); it is allocated because BFRBlender binds it.
    if (!rc) rc = vkpbrt_illumination_buffer_compile(a->accumulated);
    if (!rc) a->compiled = true;
    return rc;
}

int vkpbrt_accumulator_accumulated_illumination(vkpbrt_accumulator_t a, vkpbrt_illumination_buffer_t* out)
{
    VK_REQUIRE(a && out, "null argument");
    *out = a->accumulated;
    return VKPBRT_OK;
}

int vkpbrt_accumulator_accumulation_buffer(vkpbrt_accumulator_t a, vkpbrt_accumulation_buffer_t* out)
{
    VK_REQUIRE(a && out, "null argument");
    *out = a->acc;
    return VKPBRT_OK;
}

// Accumulator::set_camera_matrices (Accumulator.cpp:85-117)
int vkpbrt_accumulator_set_camera_matrices(vkpbrt_accumulator_t a, int frame_index, const vkpbrt_camera_matrices* cur,
                                           const vkpbrt_camera_matrices* prev)
{
    VK_REQUIRE(a && cur && prev, "null argument");
    if (a->separate_matrices) {
        if (!cur->has_proj)
            return fail(VKPBRT_ERR_MISSING_MATRICES,
                        "Accumulator::set_camera_matrices: created with separate_matrices = true, but the "
                        "CameraMatrices are missing separate matrices");
        std::memcpy(a->pc_view, cur->inv_proj, 64);
        std::memcpy(a->pc_inv_view, cur->inv_view, 64);
        if (frame_index != 0) {
            std::memcpy(a->pc_prev_view, prev->view, 64);
            float inv[16];
            vsg_inverse(prev->view, inv);                // Accumulator.cpp:100: vsg's host-side inverse, not the shader's
            a->pc_prev_pos[0] = inv[12]; a->pc_prev_pos[1] = inv[13]; a->pc_prev_pos[2] = inv[14]; a->pc_prev_pos[3] = 1.0f;
        }
    } else {
        std::memcpy(a->pc_view, cur->view, 64);
        std::memcpy(a->pc_inv_view, cur->inv_view, 64);
        if (frame_index != 0) {
            std::memcpy(a->pc_prev_view, prev->view, 64);
            // Accumulator.cpp:110-111 `prev_pos = prev.inv_view[2]; prev_pos /= prev_pos.w`: vsg's vec4 /= multiplies by the
            // reciprocal (vsg/maths/vec4.h:131-140)
            const float inv_w = 1.0f / prev->inv_view[11];
            for (int i = 0; i < 4; ++i) a->pc_prev_pos[i] = prev->inv_view[8 + i] * inv_w;
        }
    }
    a->pc_frame_number = frame_index;
    return VKPBRT_OK;
}

int vkpbrt_accumulator_set_row_range(vkpbrt_accumulator_t a, int row_begin, int row_end)
{
    VK_REQUIRE(a, "null accumulator");
    VK_REQUIRE(row_begin >= 0 && row_end <= a->height && row_begin <= row_end, "row range out of bounds");
    a->row_begin = row_begin; a->row_end = row_end;
    return VKPBRT_OK;
}

int vkpbrt_accumulator_set_force_scalar(vkpbrt_accumulator_t a, int enable)
{
    VK_REQUIRE(a, "null accumulator");
    a->force_scalar = enable != 0;
    return VKPBRT_OK;
}

int vkpbrt_accumulator_set_max_displacement_rows(vkpbrt_accumulator_t a, int rows)
{
    VK_REQUIRE(a && rows >= 0, "vkpbrt_accumulator_set_max_displacement_rows: bad argument");
    if (rows > 0 && !a->disp_violations) {
        VK_CUDA(cudaSetDevice(a->ctx->device));
        VK_CUDA(cudaMalloc((void**)&a->disp_violations, sizeof(uint32_t)));
        VK_CUDA(cudaMemsetAsync(a->disp_violations, 0, sizeof(uint32_t), joined(a->ctx)));
    }
    a->max_disp_rows = rows;
    return VKPBRT_OK;
}

int vkpbrt_accumulator_displacement_violations(vkpbrt_accumulator_t a, uint32_t* count)
{
    VK_REQUIRE(a && count, "null argument");
    *count = 0;
    if (!a->disp_violations) return VKPBRT_OK;
    VK_CUDA(cudaSetDevice(a->ctx->device));
    VK_CUDA(cudaMemcpyAsync(count, a->disp_violations, sizeof(uint32_t), cudaMemcpyDeviceToHost, joined(a->ctx)));
    VK_CUDA(cudaStreamSynchronize(joined(a->ctx)));
    return VKPBRT_OK;
}

int vkpbrt_accumulator_record(vkpbrt_accumulator_t a)
{
    VK_REQUIRE(a, "null accumulator");
    if (!a->compiled) return fail(VKPBRT_ERR_NOT_COMPILED, "Accumulator: compile_images() has not been called");
    vkpbrt_image_t src = a->original->images[0];
    vkpbrt_image_t depth = a->g->img[VKPBRT_GBUFFER_DEPTH];
    VK_REQUIRE(src->data && depth->data, "Accumulator: input images not compiled");
    vkpbrt::AccumulateParams p{};
    p.W = a->width; p.H = a->height;
    p.row_begin = a->row_begin; p.row_end = a->row_end;
    p.separate_matrices = a->separate_matrices ? 1 : 0;
    p.src_is_f16 = src->format == VKPBRT_FORMAT_R16G16B16A16_SFLOAT;
    p.frame = (uint32_t)a->pc_frame_number;
    std::memcpy(p.m_dir, a->pc_view, 64);
    std::memcpy(p.inv_view, a->pc_inv_view, 64);
    if (a->separate_matrices) {
        float proj[16];
        mat_inverse(a->pc_view, proj);               // accumulator.comp:46
        mat_mul(proj, a->pc_prev_view, p.m_prev);    // :54 proj * prevView
    } else {
        std::memcpy(p.m_prev, a->pc_prev_view, 64);
    }
    std::memcpy(p.prev_origin, a->pc_prev_pos, 16);
    const float sz[2] = {(float)a->width, (float)a->height};
    for (int i = 0; i < 2; ++i) {
        p.uv_scale[i] = sz[i] / (sz[i] - 0.5f);
        p.rcp_size[i] = 1.0f / sz[i];
    }
    for (int i = 0; i < 4; ++i) p.cur_origin[i] = a->pc_inv_view[8 + i] / a->pc_inv_view[11];
    p.src = src->data;
    p.depth = (const float*)depth->data;
    p.prev_depth = (const float*)a->acc->img[VKPBRT_ACC_PREV_DEPTH]->data;
    p.prev_illum = (const uint2*)a->acc->img[VKPBRT_ACC_PREV_ILLU]->data;
    p.prev_spp = (const uint8_t*)a->acc->img[VKPBRT_ACC_PREV_SPP]->data;
    p.motion = (uint32_t*)a->acc->img[VKPBRT_ACC_MOTION]->data;
    p.spp = (uint8_t*)a->acc->img[VKPBRT_ACC_SPP]->data;
    p.illum = (uint2*)a->accumulated->images[0]->data;
    p.depth_history = (float*)a->acc->depth_next->data;
    p.one = 1.0f; p.neg_one = -1.0f;
    p.force_scalar = a->force_scalar ? 1 : 0;
    p.max_disp_rows = a->disp_violations ? a->max_disp_rows : 0;
    p.disp_violations = a->disp_violations;
    VK_CUDA(cudaSetDevice(a->ctx->device));
    VK_CUDA(vkpbrt::launch_accumulate(p, joined(a->ctx)));
    a->ctx->launches++;
    a->acc->depth_next_valid = true;
    return VKPBRT_OK;
}

int vkpbrt_accumulator_destroy(vkpbrt_accumulator_t a)
{
    if (!a) return VKPBRT_OK;
    vkpbrt_illumination_buffer_destroy(a->accumulated);
    vkpbrt_accumulation_buffer_destroy(a->acc);
    if (a->disp_violations) cudaFree(a->disp_violations);
    delete a;
    return VKPBRT_OK;
}

// ---- shared checks for the block denoisers --------------------------------------------------------
static int check_denoiser_inputs(const char* who, uint32_t width, uint32_t height, uint32_t ww, uint32_t wh, vkpbrt_gbuffer_t g,
                                 vkpbrt_illumination_buffer_t illum, vkpbrt_accumulation_buffer_t acc)
{
    VK_REQUIRE(g && illum && acc, std::string(who) + ": null buffer");
    // denoisers/BMFR.cpp:17-22, BFR.cpp:15-20
    if (illum->type != VKPBRT_ILLUMINATION_DEMODULATED && illum->type != VKPBRT_ILLUMINATION_DEMODULATED_FLOAT)
        return fail(VKPBRT_ERR_WRONG_BUFFER_TYPE,
                    "Illumination Buffer type is required to be IlluminationBufferDemodulated/Float for BMFR");
    if (illum->images[0]->format != VKPBRT_FORMAT_R16G16B16A16_SFLOAT)
        return fail(VKPBRT_ERR_UNSUPPORTED, std::string(who) + ": the noisy input must be the accumulated rgba16f illumination "
                                                               "(Accumulator::accumulated_illumination)");
    VK_REQUIRE(ww == wh, std::string(who) + ": work_width must equal work_height");
    VK_REQUIRE(g->width == width && g->height == height, std::string(who) + ": g-buffer extent mismatch");
    VK_REQUIRE(acc->width == width && acc->height == height, std::string(who) + ": accumulation buffer extent mismatch");
    VK_REQUIRE(width >= ww && height >= wh, std::string(who) + ": image smaller than one block");
    return VKPBRT_OK;
}

// ---- BMFR ---------------------------------------------------------------------------------------
int vkpbrt_bmfr_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, uint32_t work_width, uint32_t work_height,
                       vkpbrt_gbuffer_t g, vkpbrt_illumination_buffer_t illumination, vkpbrt_accumulation_buffer_t acc,
                       uint32_t fitting_kernel, vkpbrt_bmfr_t* out)
{
    VK_REQUIRE(ctx && out, "vkpbrt_bmfr_create: null argument");
    int rc = check_denoiser_inputs("BMFR", width, height, work_width, work_height, g, illumination, acc);
    if (rc) return rc;
    const bool ok = (work_width == 32 && fitting_kernel == 256) || (work_width == 16 && fitting_kernel == 256) ||
                    (work_width == 8 && fitting_kernel == 64);
    if (!ok)
        return fail(VKPBRT_ERR_UNSUPPORTED, "BMFR: supported (work_width, fitting_kernel) are (32,256) (16,256) (8,64) "
                                            "(util/DenoiserUtils.cpp:78-124)");
    auto* b = new vkpbrt_bmfr_s();
    b->ctx = ctx; b->width = width; b->height = height; b->work = work_width; b->fitting_kernel = fitting_kernel;
    b->blocks_x = width / work_width + 2;      // BMFR.cpp:12-13, :208-209
    b->blocks_y = height / work_height + 2;
    b->g = g; b->illum = illumination; b->acc = acc;
    b->block_row_begin = 0; b->block_row_end = (int)b->blocks_y;
    if (const char* e = std::getenv("VKPBRT_BMFR_TMA")) b->tma_enabled = std::atoi(e) != 0;
    rc = make_image(ctx, VKPBRT_FORMAT_R16G16B16A16_SFLOAT, width, height, 2, &b->denoised);     // BMFR.cpp:56-75
    if (!rc) rc = make_image(ctx, VKPBRT_FORMAT_B8G8R8A8_UNORM, width, height, 1, &b->final_image);   // :78-93
    if (rc) { vkpbrt_bmfr_destroy(b); return rc; }
    *out = b;
    return VKPBRT_OK;
}

int vkpbrt_bmfr_set_debug_outputs(vkpbrt_bmfr_t b, int enable)
{
    VK_REQUIRE(b, "null bmfr");
    VK_REQUIRE(!b->compiled, "vkpbrt_bmfr_set_debug_outputs must precede compile()");
    b->debug = (enable & 1) != 0;
    b->force_generic = (enable & 2) != 0;
    return VKPBRT_OK;
}

int vkpbrt_bmfr_set_position_type(vkpbrt_bmfr_t b, int position_type)
{
    VK_REQUIRE(b && position_type >= 0 && position_type <= 2, "vkpbrt_bmfr_set_position_type: 0 (depth), 1 (world, normalised depth) or 2 (world)");
    b->position_type = position_type;
    return VKPBRT_OK;
}

int vkpbrt_bmfr_set_lane(vkpbrt_bmfr_t b, int lane)
{
    VK_REQUIRE(b && lane >= 0 && lane <= vkpbrt_context_s::kLanes, "vkpbrt_bmfr_set_lane: lane must be 0, 1 or 2");
    b->lane = lane;
    return VKPBRT_OK;
}

int vkpbrt_bmfr_compile(vkpbrt_bmfr_t b)
{
    VK_REQUIRE(b, "null bmfr");
    int rc = vkpbrt_image_compile(b->denoised);
    if (!rc) rc = vkpbrt_image_compile(b->final_image);
    if (!rc && b->debug) {
        if (!b->features) rc = make_image(b->ctx, VKPBRT_FORMAT_R16_SFLOAT, b->blocks_x * b->work, b->blocks_y * b->work, 13, &b->features);   // BMFR.cpp:96-113
        if (!rc && !b->weights) rc = make_image(b->ctx, VKPBRT_FORMAT_R32_SFLOAT, b->blocks_x, b->blocks_y, 30, &b->weights);                  // :116-133
        if (!rc) rc = vkpbrt_image_compile(b->features);
        if (!rc) rc = vkpbrt_image_compile(b->weights);
    }
    if (!rc && !b->table[0]) {
        VK_CUDA(cudaSetDevice(b->ctx->device));
        for (int i = 0; i < 2; ++i) {
            VK_CUDA(cudaMalloc((void**)&b->table[i], vkpbrt::bmfr_table_floats((int)b->work) * sizeof(float)));
            b->table_frame[i] = -1;
        }
    }
    if (!rc) b->compiled = true;
    return rc;
}

int vkpbrt_bmfr_set_block_row_range(vkpbrt_bmfr_t b, int begin, int end)
{
    VK_REQUIRE(b, "null bmfr");
    VK_REQUIRE(begin >= 0 && end <= (int)b->blocks_y && begin <= end, "block row range out of bounds");
    b->block_row_begin = begin; b->block_row_end = end;
    return VKPBRT_OK;
}

int vkpbrt_bmfr_record(vkpbrt_bmfr_t b, const vkpbrt_push_constants* pc)
{
    VK_REQUIRE(b && pc, "null argument");
    if (!b->compiled) return fail(VKPBRT_ERR_NOT_COMPILED, "BMFR: compile() has not been called");
    vkpbrt::BmfrParams p{};
    p.W = (int)b->width; p.H = (int)b->height; p.block = (int)b->work; p.fitting_kernel = (int)b->fitting_kernel;
    p.blocks_x = (int)b->blocks_x; p.blocks_y = (int)b->blocks_y;
    p.block_row_begin = b->block_row_begin; p.block_row_end = b->block_row_end;
    p.frame = pc->frame_number;
    {
        // bmfrGeneral.comp:36 / bmfrPre.comp:16: float multiply, truncation toward zero
        static const float offs[16][2] = {{.7f, .85f}, {.95f, .5f}, {.43f, .76f}, {.97f, .03f}, {.37f, .58f}, {.03f, .36f}, {.81f, .46f}, {0.f, .78f},
                                          {.36f, -.08f}, {-.06f, 0.f}, {.95f, .1f}, {.85f, .61f}, {.06f, .1f}, {.43f, .16f}, {0.f, .5f}, {.73f, .38f}};
        const float bw = (float)b->work;
        p.off_x = (int)(bw * offs[pc->frame_number % 16][0]);
        p.off_y = (int)(bw * offs[pc->frame_number % 16][1]);
    }
    p.depth = (const float*)b->g->img[VKPBRT_GBUFFER_DEPTH]->data;
    p.normal = (const float2*)b->g->img[VKPBRT_GBUFFER_NORMAL]->data;
    p.albedo = (const uchar4*)b->g->img[VKPBRT_GBUFFER_ALBEDO]->data;
    p.motion = (const uint32_t*)b->acc->img[VKPBRT_ACC_MOTION]->data;
    p.spp = (const uint8_t*)b->acc->img[VKPBRT_ACC_SPP]->data;
    p.noisy = (const uint2*)b->illum->images[0]->data;
    VK_REQUIRE(p.depth && p.normal && p.albedo && p.motion && p.spp && p.noisy, "BMFR: an input image is not compiled");
    const uint64_t layer = b->denoised->layer_pitch();
    p.denoised_prev = (const uint2*)((const char*)b->denoised->data + (uint64_t)(pc->frame_number & 1u) * layer);      // bmfrPost.comp:111
    p.denoised_next = (uint2*)((char*)b->denoised->data + (uint64_t)((pc->frame_number & 1u) ^ 1u) * layer);           // :118
    p.final_bgra = (uint32_t*)b->final_image->data;
    p.dbg_features = b->debug ? (uint16_t*)b->features->data : nullptr;
    p.dbg_weights = b->debug ? (float*)b->weights->data : nullptr;
    p.force_generic = b->force_generic ? 1 : 0;
    p.one = 1.0f; p.neg_one = -1.0f;
    VK_CUDA(cudaSetDevice(b->ctx->device));
    // the frame's block-invariant table: normally written by the previous frame's launch (its spare CTAs produce the
    // table of frame + 1); the first frame, or a frame number that does not follow the previous one, builds it here
    const int slot = (int)(pc->frame_number & 1u);
    cudaStream_t st = nullptr;
    VK_CUDA(fork_lane(b->ctx, b->lane, &st));
    if (b->table_frame[slot] != (int64_t)pc->frame_number) {
        VK_CUDA(vkpbrt::launch_bmfr_table((int)b->work, b->table[slot], pc->frame_number, st));
        b->ctx->launches++;
        b->table_frame[slot] = (int64_t)pc->frame_number;
    }
    p.table = b->table[slot];
    // The launch's extra grid row (blocks_x CTAs of fitting_kernel threads) writes the next frame's table, one element
    // per thread: that covers the work x work elements unless the image is narrower than two 32-blocks (blocks_x < 4),
    // and there is no extra row when the launch is empty.  Then the next frame builds its table itself.
    const bool builds_next = p.block_row_end > p.block_row_begin && (uint64_t)p.blocks_x * b->fitting_kernel >= (uint64_t)b->work * b->work;
    p.table_next = builds_next ? b->table[slot ^ 1] : nullptr;
    p.position_type = b->position_type;
    std::memcpy(p.inv_view, pc->view_inverse, 64);      // camParams.inverseViewMatrix / inverseProjectionMatrix (WORLD modes only)
    std::memcpy(p.inv_proj, pc->proj_inverse, 64);
    if (b->tma_enabled && b->position_type == 0) vkpbrt::bmfr_encode_tma(p);     // descriptors follow the planes bound for this frame
    VK_CUDA(vkpbrt::launch_bmfr(p, st));
    VK_CUDA(lane_done(b->ctx, b->lane));
    if (builds_next) b->table_frame[slot ^ 1] = (int64_t)(uint32_t)(pc->frame_number + 1u);
    b->ctx->launches++;
    return VKPBRT_OK;
}

int vkpbrt_bmfr_final_image(vkpbrt_bmfr_t b, vkpbrt_image_t* out)
{
    VK_REQUIRE(b && out, "null argument");
    *out = b->final_image;
    return VKPBRT_OK;
}

int vkpbrt_bmfr_image_get(vkpbrt_bmfr_t b, uint32_t which, vkpbrt_image_t* out)
{
    VK_REQUIRE(b && out, "null argument");
    switch (which) {
    case VKPBRT_BMFR_IMAGE_DENOISED: *out = b->denoised; break;
    case VKPBRT_BMFR_IMAGE_FEATURES: *out = b->features; break;
    case VKPBRT_BMFR_IMAGE_WEIGHTS: *out = b->weights; break;
    default: return fail(VKPBRT_ERR_INVALID_ARGUMENT, "vkpbrt_bmfr_image_get: unknown image");
    }
    if (!*out) return fail(VKPBRT_ERR_INVALID_ARGUMENT, "vkpbrt_bmfr_image_get: debug outputs are not enabled");
    return VKPBRT_OK;
}

int vkpbrt_bmfr_destroy(vkpbrt_bmfr_t b)
{
    if (!b) return VKPBRT_OK;
    vkpbrt_image_release(b->denoised);
    vkpbrt_image_release(b->final_image);
    vkpbrt_image_release(b->features);
    vkpbrt_image_release(b->weights);
    for (int i = 0; i < 2; ++i)
        if (b->table[i]) cudaFree(b->table[i]);
    delete b;
    return VKPBRT_OK;
}

// ---- BFR ----------------------------------------------------------------------------------------
int vkpbrt_bfr_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, uint32_t work_width, uint32_t work_height,
                      vkpbrt_gbuffer_t g, vkpbrt_illumination_buffer_t illumination, vkpbrt_accumulation_buffer_t acc,
                      vkpbrt_bfr_t* out)
{
    VK_REQUIRE(ctx && out, "vkpbrt_bfr_create: null argument");
    int rc = check_denoiser_inputs("BFR", width, height, work_width, work_height, g, illumination, acc);
    if (rc) return rc;
    if (work_width != 8 && work_width != 16 && work_width != 32)
        return fail(VKPBRT_ERR_UNSUPPORTED, "BFR: supported block sizes are 8, 16, 32 (util/DenoiserUtils.cpp:22-70)");
    auto* b = new vkpbrt_bfr_s();
    b->ctx = ctx; b->width = width; b->height = height; b->work = work_width;
    b->blocks_x = width / work_width + 2;      // BFR.cpp:134
    b->blocks_y = height / work_height + 2;
    b->g = g; b->illum = illumination; b->acc = acc;
    for (int i = 0; i < 40; ++i) {              // bfr.comp:134, t = i + 1
        // evaluated in double and rounded once: independent of libm's float variants / constant folding
        const float t = (float)(i + 1);
        b->lr_exp[i] = (float)std::exp((double)(-.116f * t));
        b->lr_sqrt[i] = sqrtf(1 - (float)std::pow((double).7314f, (double)(i + 1)));
        b->lr_den[i] = 1 - (float)std::pow((double).3f, (double)(i + 1));
    }
    rc = make_image(ctx, VKPBRT_FORMAT_R16G16B16A16_SFLOAT, width, height, 2, &b->denoised);
    if (!rc) rc = make_image(ctx, VKPBRT_FORMAT_B8G8R8A8_UNORM, width, height, 1, &b->final_image);
    if (rc) { vkpbrt_bfr_destroy(b); return rc; }
    *out = b;
    return VKPBRT_OK;
}

int vkpbrt_bfr_set_lane(vkpbrt_bfr_t b, int lane)
{
    VK_REQUIRE(b && lane >= 0 && lane <= vkpbrt_context_s::kLanes, "vkpbrt_bfr_set_lane: lane must be 0, 1 or 2");
    b->lane = lane;
    return VKPBRT_OK;
}

int vkpbrt_bfr_compile(vkpbrt_bfr_t b)
{
    VK_REQUIRE(b, "null bfr");
    int rc = vkpbrt_image_compile(b->denoised);
    if (!rc) rc = vkpbrt_image_compile(b->final_image);
    if (!rc) b->compiled = true;
    return rc;
}

int vkpbrt_bfr_record(vkpbrt_bfr_t b, const vkpbrt_push_constants* pc)
{
    VK_REQUIRE(b && pc, "null argument");
    if (!b->compiled) return fail(VKPBRT_ERR_NOT_COMPILED, "BFR: compile() has not been called");
    vkpbrt::BfrParams p{};
    p.W = (int)b->width; p.H = (int)b->height; p.block = (int)b->work;
    p.blocks_x = (int)b->blocks_x; p.blocks_y = (int)b->blocks_y;
    p.frame = pc->frame_number;
    std::memcpy(p.lr_exp, b->lr_exp, sizeof p.lr_exp);
    std::memcpy(p.lr_sqrt, b->lr_sqrt, sizeof p.lr_sqrt);
    std::memcpy(p.lr_den, b->lr_den, sizeof p.lr_den);
    p.depth = (const float*)b->g->img[VKPBRT_GBUFFER_DEPTH]->data;
    p.normal = (const float2*)b->g->img[VKPBRT_GBUFFER_NORMAL]->data;
    p.albedo = (const uchar4*)b->g->img[VKPBRT_GBUFFER_ALBEDO]->data;
    p.motion = (const uint32_t*)b->acc->img[VKPBRT_ACC_MOTION]->data;
    p.spp = (const uint8_t*)b->acc->img[VKPBRT_ACC_SPP]->data;
    p.noisy = (const uint2*)b->illum->images[0]->data;
    VK_REQUIRE(p.depth && p.normal && p.albedo && p.motion && p.spp && p.noisy, "BFR: an input image is not compiled");
    const uint64_t layer = b->denoised->layer_pitch();
    p.denoised_prev = (const uint2*)((const char*)b->denoised->data + (uint64_t)(pc->frame_number & 1u) * layer);
    p.denoised_next = (uint2*)((char*)b->denoised->data + (uint64_t)((pc->frame_number & 1u) ^ 1u) * layer);
    p.final_bgra = (uint32_t*)b->final_image->data;
    VK_CUDA(cudaSetDevice(b->ctx->device));
    cudaStream_t st = nullptr;
    VK_CUDA(fork_lane(b->ctx, b->lane, &st));
    VK_CUDA(vkpbrt::launch_bfr(p, st));
    VK_CUDA(lane_done(b->ctx, b->lane));
    b->ctx->launches++;
    return VKPBRT_OK;
}

int vkpbrt_bfr_final_image(vkpbrt_bfr_t b, vkpbrt_image_t* out)
{
    VK_REQUIRE(b && out, "null argument");
    *out = b->final_image;
    return VKPBRT_OK;
}

int vkpbrt_bfr_denoised_image(vkpbrt_bfr_t b, vkpbrt_image_t* out)
{
    VK_REQUIRE(b && out, "null argument");
    *out = b->denoised;
    return VKPBRT_OK;
}

int vkpbrt_bfr_destroy(vkpbrt_bfr_t b)
{
    if (!b) return VKPBRT_OK;
    vkpbrt_image_release(b->denoised);
    vkpbrt_image_release(b->final_image);
    delete b;
    return VKPBRT_OK;
}

// ---- BFRBlender -----------------------------------------------------------------------------------
int vkpbrt_bfr_blender_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, vkpbrt_image_t average,
                              vkpbrt_image_t average_squared, vkpbrt_image_t denoised0, vkpbrt_image_t denoised1,
                              vkpbrt_image_t denoised2, uint32_t work_width, uint32_t work_height, uint32_t filter_radius,
                              vkpbrt_bfr_blender_t* out)
{
    VK_REQUIRE(ctx && out && average && average_squared && denoised0 && denoised1 && denoised2, "vkpbrt_bfr_blender_create: null argument");
    VK_REQUIRE(average->format == VKPBRT_FORMAT_R16G16B16A16_SFLOAT && average_squared->format == VKPBRT_FORMAT_R16G16B16A16_SFLOAT,
               "BFRBlender: average images must be rgba16f (bfrBlender.comp:4-5)");
    for (vkpbrt_image_t d : {denoised0, denoised1, denoised2})
        VK_REQUIRE(d->format == VKPBRT_FORMAT_B8G8R8A8_UNORM && same_extent(d, width, height), "BFRBlender: denoised images must be BGRA8 of the frame size");
    VK_REQUIRE(same_extent(average, width, height) && same_extent(average_squared, width, height), "BFRBlender: extent mismatch");
    auto* b = new vkpbrt_bfr_blender_s();
    b->ctx = ctx; b->width = width; b->height = height; b->work_width = work_width; b->work_height = work_height;
    b->radius = filter_radius;
    b->average = average; b->average_squared = average_squared;
    b->den[0] = denoised0; b->den[1] = denoised1; b->den[2] = denoised2;
    int rc = make_image(ctx, VKPBRT_FORMAT_B8G8R8A8_UNORM, width, height, 1, &b->final_image);
    if (rc) { delete b; return rc; }
    *out = b;
    return VKPBRT_OK;
}

int vkpbrt_bfr_blender_compile(vkpbrt_bfr_blender_t b)
{
    VK_REQUIRE(b, "null blender");
    int rc = vkpbrt_image_compile(b->final_image);
    if (!rc) b->compiled = true;
    return rc;
}

int vkpbrt_bfr_blender_record(vkpbrt_bfr_blender_t b)
{
    VK_REQUIRE(b, "null blender");
    if (!b->compiled) return fail(VKPBRT_ERR_NOT_COMPILED, "BFRBlender: compile() has not been called");
    vkpbrt::BlendParams p{};
    p.W = (int)b->width; p.H = (int)b->height; p.radius = (int)b->radius;
    p.average = (const uint2*)b->average->data;
    p.average_squared = (const uint2*)b->average_squared->data;
    p.denoised0 = (const uint32_t*)b->den[0]->data;
    p.denoised1 = (const uint32_t*)b->den[1]->data;
    p.denoised2 = (const uint32_t*)b->den[2]->data;
    p.final_bgra = (uint32_t*)b->final_image->data;
    p.one = 1.0f; p.neg_one = -1.0f;
    VK_REQUIRE(p.average && p.average_squared && p.denoised0 && p.denoised1 && p.denoised2, "BFRBlender: an input image is not compiled");
    VK_CUDA(cudaSetDevice(b->ctx->device));
    VK_CUDA(vkpbrt::launch_bfr_blend(p, joined(b->ctx)));
    b->ctx->launches++;
    return VKPBRT_OK;
}

int vkpbrt_bfr_blender_final_image(vkpbrt_bfr_blender_t b, vkpbrt_image_t* out)
{
    VK_REQUIRE(b && out, "null argument");
    *out = b->final_image;
    return VKPBRT_OK;
}

int vkpbrt_bfr_blender_destroy(vkpbrt_bfr_blender_t b)
{
    if (!b) return VKPBRT_OK;
    vkpbrt_image_release(b->final_image);
    delete b;
    return VKPBRT_OK;
}

// ---- Taa ------------------------------------------------------------------------------------------
int vkpbrt_taa_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, uint32_t work_width, uint32_t work_height,
                      vkpbrt_gbuffer_t g, vkpbrt_accumulation_buffer_t acc, vkpbrt_image_t denoised, vkpbrt_taa_t* out)
{
    (void)g;   // Taa.cpp:4-80 takes the g-buffer but binds nothing from it
    VK_REQUIRE(ctx && out && acc && denoised, "vkpbrt_taa_create: null argument");
    VK_REQUIRE(denoised->format == VKPBRT_FORMAT_B8G8R8A8_UNORM, "Taa: the denoised image must be BGRA8 (denoiser final)");
    VK_REQUIRE(same_extent(denoised, width, height) && acc->width == width && acc->height == height, "Taa: extent mismatch");
    auto* t = new vkpbrt_taa_s();
    t->ctx = ctx; t->width = width; t->height = height; t->work_width = work_width; t->work_height = work_height;
    t->acc = acc; t->denoised = denoised;
    t->row_begin = 0; t->row_end = (int)height;
    // handles only; both views alias the ping-pong pair allocated in compile()
    int rc = vkpbrt_image_wrap(ctx, VKPBRT_FORMAT_B8G8R8A8_UNORM, width, height, 1, nullptr, &t->final_image);   // Taa.cpp:42
    if (!rc) rc = vkpbrt_image_wrap(ctx, VKPBRT_FORMAT_R8G8B8A8_UNORM, width, height, 1, nullptr, &t->history);  // Taa.cpp:24
    if (rc) { vkpbrt_taa_destroy(t); return rc; }
    *out = t;
    return VKPBRT_OK;
}

int vkpbrt_taa_set_fix_swizzle(vkpbrt_taa_t t, int fix)
{
    VK_REQUIRE(t, "null taa");
    t->fix_swizzle = fix != 0;
    return VKPBRT_OK;
}

int vkpbrt_taa_set_force_scalar(vkpbrt_taa_t t, int enable)
{
    VK_REQUIRE(t, "null taa");
    t->force_scalar = enable ? 1 : 0;
    return VKPBRT_OK;
}

int vkpbrt_taa_set_strip_rows(vkpbrt_taa_t t, int rows)
{
    VK_REQUIRE(t && rows >= 0 && rows <= 4096, "vkpbrt_taa_set_strip_rows: 0 (automatic) or a strip height in rows");
    t->strip_rows = rows;
    return VKPBRT_OK;
}

int vkpbrt_taa_compile(vkpbrt_taa_t t)
{
    VK_REQUIRE(t, "null taa");
    if (t->compiled) return VKPBRT_OK;
    const size_t bytes = (size_t)t->width * t->height * 4;
    VK_CUDA(cudaSetDevice(t->ctx->device));
    for (auto& b : t->buf) {
        VK_CUDA(cudaMalloc(&b, bytes));
        VK_CUDA(cudaMemsetAsync(b, 0, bytes, joined(t->ctx)));
    }
    t->final_image->data = t->buf[1];
    t->history->data = t->buf[0];
    t->compiled = true;
    return VKPBRT_OK;
}

int vkpbrt_taa_set_row_range(vkpbrt_taa_t t, int row_begin, int row_end)
{
    VK_REQUIRE(t, "null taa");
    VK_REQUIRE(row_begin >= 0 && row_end <= (int)t->height && row_begin <= row_end, "row range out of bounds");
    t->row_begin = row_begin; t->row_end = row_end;
    return VKPBRT_OK;
}

int vkpbrt_taa_record_parts(vkpbrt_taa_t t, const vkpbrt_push_constants* pc, int row_begin, int row_end, int row_begin2, int row_end2, int last)
{
    VK_REQUIRE(t && pc, "null argument");
    if (!t->compiled) return fail(VKPBRT_ERR_NOT_COMPILED, "Taa: compile() has not been called");
    VK_REQUIRE(row_begin >= 0 && row_end <= (int)t->height && row_begin <= row_end, "Taa: row range out of bounds");
    VK_REQUIRE(row_begin2 >= 0 && row_end2 <= (int)t->height && row_begin2 <= row_end2, "Taa: second row range out of bounds");
    VK_REQUIRE(row_end2 == row_begin2 || row_end == row_begin || row_end <= row_begin2 || row_end2 <= row_begin, "Taa: the two row ranges overlap");
    vkpbrt::TaaParams p{};
    p.W = (int)t->width; p.H = (int)t->height;
    p.row_begin = row_begin; p.row_end = row_end;
    p.row_begin2 = row_begin2; p.row_end2 = row_end2;
    p.frame = pc->frame_number;
    p.fix_swizzle = t->fix_swizzle;
    p.motion = (const uint32_t*)t->acc->img[VKPBRT_ACC_MOTION]->data;
    p.denoised = (const uint32_t*)t->denoised->data;
    void* in = t->history->data;
    void* outb = (in == t->buf[0]) ? t->buf[1] : t->buf[0];
    p.history = (const uint32_t*)in;
    p.final_bgra = (uint32_t*)outb;
    VK_REQUIRE(p.motion && p.denoised, "Taa: an input image is not compiled");
    p.one = 1.0f; p.neg_one = -1.0f;
    p.force_scalar = t->force_scalar;
    p.rows_per_warp = t->strip_rows;
    VK_CUDA(cudaSetDevice(t->ctx->device));
    VK_CUDA(vkpbrt::launch_taa(p, joined(t->ctx)));
    if (row_end > row_begin || row_end2 > row_begin2) t->ctx->launches++;
    if (last) {
        // Taa.cpp:106 copy final -> history: both handles now view this frame's output
        t->final_image->data = outb;
        t->history->data = outb;
    }
    return VKPBRT_OK;
}

int vkpbrt_taa_record_part(vkpbrt_taa_t t, const vkpbrt_push_constants* pc, int row_begin, int row_end, int last)
{
    return vkpbrt_taa_record_parts(t, pc, row_begin, row_end, 0, 0, last);
}

int vkpbrt_taa_record(vkpbrt_taa_t t, const vkpbrt_push_constants* pc)
{
    VK_REQUIRE(t, "null argument");
    return vkpbrt_taa_record_part(t, pc, t->row_begin, t->row_end, 1);
}

int vkpbrt_taa_final_image(vkpbrt_taa_t t, vkpbrt_image_t* out)
{
    VK_REQUIRE(t && out, "null argument");
    *out = t->final_image;
    return VKPBRT_OK;
}

int vkpbrt_taa_history_image(vkpbrt_taa_t t, vkpbrt_image_t* out)
{
    VK_REQUIRE(t && out, "null argument");
    *out = t->history;
    return VKPBRT_OK;
}

int vkpbrt_taa_destroy(vkpbrt_taa_t t)
{
    if (!t) return VKPBRT_OK;
    for (auto& b : t->buf)
        if (b) cudaFree(b);
    vkpbrt_image_release(t->final_image);
    vkpbrt_image_release(t->history);
    delete t;
    return VKPBRT_OK;
}

// ---- FormatConverter (source/renderModules/FormatConverter.cpp:4-93, shaders/formatConverter.comp) ---
int vkpbrt_format_converter_create(vkpbrt_context_t ctx, vkpbrt_image_t src_image, uint32_t dst_format, uint32_t work_width,
                                   uint32_t work_height, vkpbrt_format_converter_t* out)
{
    VK_REQUIRE(ctx && src_image && out, "vkpbrt_format_converter_create: null argument");
    if (dst_format != VKPBRT_FORMAT_B8G8R8A8_UNORM)
        return fail(VKPBRT_ERR_UNSUPPORTED, "FormatConverter::Unknown format");            // FormatConverter.cpp:13-19
    if (src_image->format != VKPBRT_FORMAT_R32G32B32A32_SFLOAT && src_image->format != VKPBRT_FORMAT_R16G16B16A16_SFLOAT &&
        src_image->format != VKPBRT_FORMAT_R8G8B8A8_UNORM)
        return fail(VKPBRT_ERR_UNSUPPORTED, "FormatConverter: the source must be a four-channel image (rgba32f, rgba16f or rgba8)");
    auto* f = new vkpbrt_format_converter_s();
    f->ctx = ctx; f->src = src_image; f->work_width = work_width; f->work_height = work_height;
    int rc = make_image(ctx, VKPBRT_FORMAT_B8G8R8A8_UNORM, src_image->width, src_image->height, 1, &f->final_image);   // :38-47
    if (rc) { delete f; return rc; }
    *out = f;
    return VKPBRT_OK;
}

int vkpbrt_format_converter_compile_images(vkpbrt_format_converter_t f)
{
    VK_REQUIRE(f, "null format converter");
    int rc = vkpbrt_image_compile(f->final_image);
    if (!rc) f->compiled = true;
    return rc;
}

int vkpbrt_format_converter_record(vkpbrt_format_converter_t f)
{
    VK_REQUIRE(f, "null format converter");
    if (!f->compiled) return fail(VKPBRT_ERR_NOT_COMPILED, "FormatConverter: compile_images() has not been called");
    VK_REQUIRE(f->src->data, "FormatConverter: the source image is not compiled");
    vkpbrt::FormatConvertParams p{};
    p.W = (int)f->src->width; p.H = (int)f->src->height;
    p.src_format = f->src->format == VKPBRT_FORMAT_R32G32B32A32_SFLOAT ? 0 : (f->src->format == VKPBRT_FORMAT_R16G16B16A16_SFLOAT ? 1 : 2);
    p.src = f->src->data;
    p.dst_bgra = (uint32_t*)f->final_image->data;
    VK_CUDA(cudaSetDevice(f->ctx->device));
    VK_CUDA(vkpbrt::launch_format_convert(p, joined(f->ctx)));
    f->ctx->launches++;
    return VKPBRT_OK;
}

int vkpbrt_format_converter_final_image(vkpbrt_format_converter_t f, vkpbrt_image_t* out)
{
    VK_REQUIRE(f && out, "null argument");
    *out = f->final_image;
    return VKPBRT_OK;
}

int vkpbrt_format_converter_destroy(vkpbrt_format_converter_t f)
{
    if (!f) return VKPBRT_OK;
    vkpbrt_image_release(f->final_image);
    delete f;
    return VKPBRT_OK;
}

// ---- offline sequences: the import conversions of GBufferIO on the device (source/io/RenderIO.cpp:101-120, :160-195) ----
int vkpbrt_gbuffer_import_record(vkpbrt_gbuffer_t g, vkpbrt_image_t position, const float* inv_view, vkpbrt_image_t normal,
                                 vkpbrt_image_t albedo)
{
    VK_REQUIRE(g, "vkpbrt_gbuffer_import_record: null g-buffer");
    VK_REQUIRE(!position || inv_view, "vkpbrt_gbuffer_import_record: a position plane needs the frame's inverse view matrix");
    for (vkpbrt_image_t i : {position, normal, albedo})
        VK_REQUIRE(!i || (i->format == VKPBRT_FORMAT_R32G32B32A32_SFLOAT && same_extent(i, g->width, g->height) && i->data),
                   "vkpbrt_gbuffer_import_record: planes are compiled rgba32f images of the g-buffer's size");
    vkpbrt::GBufferImportParams p{};
    p.W = (int)g->width; p.H = (int)g->height;
    if (position) {
        // RenderIO.cpp:109-110 `camera_pos = inv_view[2]; camera_pos /= camera_pos.w`: vsg's vec4 /= multiplies by the
        // reciprocal (vsg/maths/vec4.h:131-140) -- not the same rounding as a division
        const float inv_w = 1.0f / inv_view[11];
        for (int i = 0; i < 3; ++i) p.camera[i] = inv_view[8 + i] * inv_w;
        p.position = (const float4*)position->data;
        p.depth = (float*)g->img[VKPBRT_GBUFFER_DEPTH]->data;
        VK_REQUIRE(p.depth, "vkpbrt_gbuffer_import_record: the g-buffer is not compiled");
    }
    if (normal) {
        p.normal = (const float4*)normal->data;
        p.normal_out = (float2*)g->img[VKPBRT_GBUFFER_NORMAL]->data;
        VK_REQUIRE(p.normal_out, "vkpbrt_gbuffer_import_record: the g-buffer is not compiled");
    }
    if (albedo) {
        p.albedo = (const float4*)albedo->data;
        p.albedo_out = (uint32_t*)g->img[VKPBRT_GBUFFER_ALBEDO]->data;
        VK_REQUIRE(p.albedo_out, "vkpbrt_gbuffer_import_record: the g-buffer is not compiled");
    }
    VK_CUDA(cudaSetDevice(g->ctx->device));
    VK_CUDA(vkpbrt::launch_gbuffer_import(p, joined(g->ctx)));
    g->ctx->launches++;
    return VKPBRT_OK;
}

// ---- producer side: demodulated illumination (shaders/ptRaygen.rgen:81-88) ---------------------------
int vkpbrt_demodulate_record(vkpbrt_context_t ctx, vkpbrt_image_t radiance, vkpbrt_image_t albedo, vkpbrt_image_t position_x,
                             vkpbrt_image_t demodulated)
{
    VK_REQUIRE(ctx && radiance && albedo && position_x && demodulated, "vkpbrt_demodulate_record: null argument");
    VK_REQUIRE(radiance->format == VKPBRT_FORMAT_R32G32B32A32_SFLOAT && albedo->format == VKPBRT_FORMAT_R32G32B32A32_SFLOAT &&
               demodulated->format == VKPBRT_FORMAT_R32G32B32A32_SFLOAT && position_x->format == VKPBRT_FORMAT_R32_SFLOAT,
               "vkpbrt_demodulate_record: radiance / albedo / output are rgba32f, position_x is r32f");
    VK_REQUIRE(same_extent(albedo, radiance->width, radiance->height) && same_extent(position_x, radiance->width, radiance->height) &&
               same_extent(demodulated, radiance->width, radiance->height), "vkpbrt_demodulate_record: extent mismatch");
    VK_REQUIRE(radiance->data && albedo->data && position_x->data && demodulated->data, "vkpbrt_demodulate_record: an image is not compiled");
    vkpbrt::DemodulateParams p{};
    p.W = (int)radiance->width; p.H = (int)radiance->height;
    p.radiance = (const float4*)radiance->data;
    p.albedo = (const float4*)albedo->data;
    p.position_x = (const float*)position_x->data;
    p.out = (float4*)demodulated->data;
    VK_CUDA(cudaSetDevice(ctx->device));
    VK_CUDA(vkpbrt::launch_demodulate(p, joined(ctx)));
    ctx->launches++;
    return VKPBRT_OK;
}

// ---- Vulkan interop --------------------------------------------------------------------------------
int vkpbrt_import_external_memory_fd_ex(vkpbrt_context_t ctx, int fd, uint64_t allocation_size, uint64_t offset, uint64_t size,
                                        int dedicated, vkpbrt_external_memory_t* out, void** device_ptr)
{
    VK_REQUIRE(ctx && out && device_ptr, "null argument");
    VK_REQUIRE(fd >= 0, "vkpbrt_import_external_memory_fd: invalid file descriptor");
    VK_REQUIRE(size > 0 && offset <= allocation_size && size <= allocation_size - offset, "mapped range exceeds the allocation");
    VK_CUDA(cudaSetDevice(ctx->device));
    cudaExternalMemoryHandleDesc hd{};
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    hd.size = allocation_size;
    hd.flags = dedicated ? cudaExternalMemoryDedicated : 0;
    cudaExternalMemory_t mem;
    VK_CUDA(cudaImportExternalMemory(&mem, &hd));      // on success the fd belongs to CUDA (the caller must not close it)
    cudaExternalMemoryBufferDesc bd{};
    bd.offset = offset;
    bd.size = size;
    void* ptr = nullptr;
    cudaError_t e = cudaExternalMemoryGetMappedBuffer(&ptr, mem, &bd);
    if (e != cudaSuccess) { cudaDestroyExternalMemory(mem); return fail_cuda(e, "cudaExternalMemoryGetMappedBuffer"); }
    *out = new vkpbrt_external_memory_s{ctx, mem, ptr};
    *device_ptr = ptr;
    return VKPBRT_OK;
}

int vkpbrt_import_external_memory_fd(vkpbrt_context_t ctx, int fd, uint64_t allocation_size, uint64_t offset, uint64_t size,
                                     vkpbrt_external_memory_t* out, void** device_ptr)
{
    return vkpbrt_import_external_memory_fd_ex(ctx, fd, allocation_size, offset, size, 0, out, device_ptr);
}

int vkpbrt_external_memory_destroy(vkpbrt_external_memory_t m)
{
    if (!m) return VKPBRT_OK;
    cudaFree(m->ptr);
    cudaDestroyExternalMemory(m->mem);
    delete m;
    return VKPBRT_OK;
}

int vkpbrt_import_external_semaphore_fd(vkpbrt_context_t ctx, int fd, int timeline, vkpbrt_external_semaphore_t* out)
{
    VK_REQUIRE(ctx && out, "null argument");
    VK_CUDA(cudaSetDevice(ctx->device));
    cudaExternalSemaphoreHandleDesc hd{};
    hd.type = timeline ? cudaExternalSemaphoreHandleTypeTimelineSemaphoreFd : cudaExternalSemaphoreHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    cudaExternalSemaphore_t sem;
    VK_CUDA(cudaImportExternalSemaphore(&sem, &hd));
    *out = new vkpbrt_external_semaphore_s{ctx, sem, timeline != 0};
    return VKPBRT_OK;
}

int vkpbrt_external_semaphore_wait(vkpbrt_external_semaphore_t s, uint64_t value)
{
    VK_REQUIRE(s, "null semaphore");
    cudaExternalSemaphoreWaitParams wp{};
    wp.params.fence.value = value;
    VK_CUDA(cudaWaitExternalSemaphoresAsync(&s->sem, &wp, 1, joined(s->ctx)));
    return VKPBRT_OK;
}

int vkpbrt_external_semaphore_signal(vkpbrt_external_semaphore_t s, uint64_t value)
{
    VK_REQUIRE(s, "null semaphore");
    cudaExternalSemaphoreSignalParams sp{};
    sp.params.fence.value = value;
    VK_CUDA(cudaSignalExternalSemaphoresAsync(&s->sem, &sp, 1, joined(s->ctx)));
    return VKPBRT_OK;
}

int vkpbrt_external_semaphore_destroy(vkpbrt_external_semaphore_t s)
{
    if (!s) return VKPBRT_OK;
    cudaDestroyExternalSemaphore(s->sem);
    delete s;
    return VKPBRT_OK;
}

// ---- band-sharded runs: NVLink peer memory -----------------------------------------------------------
using vkpbrt::HaloCopy;
using vkpbrt::HaloPushParams;
using vkpbrt::HaloWaitParams;
using vkpbrt::kHaloMaxPeers;
static_assert(sizeof(vkpbrt_halo_copy) == sizeof(HaloCopy) && VKPBRT_HALO_MAX_PEERS == kHaloMaxPeers, "halo table layout");
static_assert(sizeof(cudaIpcMemHandle_t) == VKPBRT_PEER_HANDLE_BYTES, "IPC handle size");

int vkpbrt_peer_export(vkpbrt_context_t ctx, const void* device_ptr, uint8_t handle[VKPBRT_PEER_HANDLE_BYTES], uint64_t* offset)
{
    VK_REQUIRE(ctx && device_ptr && handle && offset, "null argument");
    VK_CUDA(cudaSetDevice(ctx->device));
#ifdef VKPBRT_HOSTSIM
    unsigned long long base = (unsigned long long)(uintptr_t)device_ptr;     // test emulator: one address space
#else
    // the handle names the whole allocation: find its base through the driver entry point the runtime already holds
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    VK_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    VK_REQUIRE(fn && qr == cudaDriverEntryPointSuccess, "cuMemGetAddressRange is not available");
    unsigned long long base = 0;
    size_t size = 0;
    const int rc = reinterpret_cast<range_fn>(fn)(&base, &size, (unsigned long long)(uintptr_t)device_ptr);
    VK_REQUIRE(rc == 0 && base != 0, "not a device allocation");
#endif
    cudaIpcMemHandle_t h;
    VK_CUDA(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base)));
    memcpy(handle, &h, sizeof(h));
    *offset = (uint64_t)((uintptr_t)device_ptr - (uintptr_t)base);
    return VKPBRT_OK;
}

int vkpbrt_peer_open(vkpbrt_context_t ctx, const uint8_t handle[VKPBRT_PEER_HANDLE_BYTES], void** base)
{
    VK_REQUIRE(ctx && handle && base, "null argument");
    VK_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    VK_CUDA(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
    return VKPBRT_OK;
}

int vkpbrt_peer_close(vkpbrt_context_t ctx, void* base)
{
    VK_REQUIRE(ctx, "null context");
    if (!base) return VKPBRT_OK;
    VK_CUDA(cudaSetDevice(ctx->device));
    VK_CUDA(cudaIpcCloseMemHandle(base));
    return VKPBRT_OK;
}

struct vkpbrt_halo_exchange_s {
    vkpbrt_context_t ctx = nullptr;
    HaloPushParams push{};
    HaloWaitParams wait{};
    bool has_start = false;
    void* device_block = nullptr;      // [copy table][counter u32, error u32][gate_ns u64 x2][wait_ns u64 x2]
    cudaEvent_t ordered = nullptr;
};

int vkpbrt_halo_exchange_create(vkpbrt_context_t ctx, const vkpbrt_halo_exchange_desc* d, uint32_t timeout_ms, vkpbrt_halo_exchange_t* out)
{
    VK_REQUIRE(ctx && d && out, "null argument");
    VK_REQUIRE(d->n_copies == 0 || d->copies, "null copy table");
    VK_REQUIRE(d->n_announce <= (uint32_t)kHaloMaxPeers && d->n_ready <= (uint32_t)kHaloMaxPeers &&
                   d->n_done <= (uint32_t)kHaloMaxPeers && d->n_wait <= (uint32_t)kHaloMaxPeers, "too many peers");
    VK_REQUIRE((d->n_announce == 0 || d->announce_flags) && (d->n_ready == 0 || d->ready_flags) &&
                   (d->n_done == 0 || d->done_flags) && (d->n_wait == 0 || d->wait_flags), "null flag list");
    VK_CUDA(cudaSetDevice(ctx->device));
    auto* x = new vkpbrt_halo_exchange_s();
    x->ctx = ctx;
    const size_t table_bytes = (size_t)d->n_copies * sizeof(HaloCopy);
    const size_t tail = (table_bytes + 15) / 16 * 16;
    cudaError_t e = cudaMalloc(&x->device_block, tail + 48);
    if (e == cudaSuccess) e = cudaMemset(x->device_block, 0, tail + 48);
    if (e == cudaSuccess && table_bytes) e = cudaMemcpy(x->device_block, d->copies, table_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&x->ordered, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        if (x->device_block) cudaFree(x->device_block);
        delete x;
        return fail_cuda(e, "vkpbrt_halo_exchange_create");
    }
    auto* base = static_cast<unsigned char*>(x->device_block);
    HaloPushParams& p = x->push;
    p.copies = reinterpret_cast<const HaloCopy*>(base);
    p.n_copies = (int)d->n_copies;
    p.n_announce = (int)d->n_announce;
    p.n_ready = (int)d->n_ready;
    p.n_done = (int)d->n_done;
    for (uint32_t i = 0; i < d->n_announce; ++i) p.announce_flags[i] = d->announce_flags[i];
    for (uint32_t i = 0; i < d->n_ready; ++i) p.ready_flags[i] = d->ready_flags[i];
    for (uint32_t i = 0; i < d->n_done; ++i) p.done_flags[i] = d->done_flags[i];
    p.counter = reinterpret_cast<uint32_t*>(base + tail);
    p.error = p.counter + 1;
    p.gate_ns = reinterpret_cast<unsigned long long*>(base + tail + 16);
    p.timeout_ns = (unsigned long long)timeout_ms * 1000000ull;
    x->has_start = d->n_copies || d->n_announce || d->n_done;
    HaloWaitParams& w = x->wait;
    w.n = (int)d->n_wait;
    for (uint32_t i = 0; i < d->n_wait; ++i) w.flags[i] = d->wait_flags[i];
    w.error = p.error;
    w.wait_ns = reinterpret_cast<unsigned long long*>(base + tail + 32);
    w.timeout_ns = p.timeout_ns;
    *out = x;
    return VKPBRT_OK;
}

int vkpbrt_halo_exchange_start_gated(vkpbrt_halo_exchange_t x, void* comm_stream, void* after_stream, uint32_t value, uint32_t gate_value)
{
    VK_REQUIRE(x, "null exchange");
    if (!x->has_start) return VKPBRT_OK;
    cudaStream_t comm = comm_stream ? (cudaStream_t)comm_stream : joined(x->ctx);
    cudaStream_t after = after_stream ? (cudaStream_t)after_stream : joined(x->ctx);
    if (comm != after) {
        VK_CUDA(cudaEventRecord(x->ordered, after));
        VK_CUDA(cudaStreamWaitEvent(comm, x->ordered, 0));
    }
    x->push.value = value;
    x->push.gate_value = gate_value;
    // 16 CTAs x 256 threads x 16 B x 4 in flight = 256 KB per sweep of one block of rows
    VK_CUDA(vkpbrt::launch_halo_push(x->push, x->push.n_copies ? 16 : 1, comm));
    x->ctx->launches++;
    return VKPBRT_OK;
}

int vkpbrt_halo_exchange_start(vkpbrt_halo_exchange_t x, void* comm_stream, void* after_stream, uint32_t value)
{
    return vkpbrt_halo_exchange_start_gated(x, comm_stream, after_stream, value, value);
}

int vkpbrt_halo_exchange_wait(vkpbrt_halo_exchange_t x, void* stream, uint32_t value)
{
    VK_REQUIRE(x, "null exchange");
    if (x->wait.n == 0) return VKPBRT_OK;
    x->wait.value = value;
    VK_CUDA(vkpbrt::launch_halo_wait(x->wait, stream ? (cudaStream_t)stream : joined(x->ctx)));
    x->ctx->launches++;
    return VKPBRT_OK;
}

int vkpbrt_halo_exchange_stats(vkpbrt_halo_exchange_t x, uint64_t* gate_ns, uint64_t* wait_ns, uint32_t* error)
{
    VK_REQUIRE(x, "null exchange");
    VK_CUDA(cudaSetDevice(x->ctx->device));
    VK_CUDA(cudaDeviceSynchronize());
    unsigned long long host[4] = {0, 0, 0, 0};
    uint32_t words[2] = {0, 0};
    VK_CUDA(cudaMemcpy(host, x->push.gate_ns, sizeof(host), cudaMemcpyDeviceToHost));
    VK_CUDA(cudaMemcpy(words, x->push.counter, sizeof(words), cudaMemcpyDeviceToHost));
    if (gate_ns) *gate_ns = host[0];
    if (wait_ns) *wait_ns = host[2];
    if (error) *error = words[1];
    return VKPBRT_OK;
}

int vkpbrt_halo_exchange_destroy(vkpbrt_halo_exchange_t x)
{
    if (!x) return VKPBRT_OK;
    if (x->ordered) cudaEventDestroy(x->ordered);
    if (x->device_block) cudaFree(x->device_block);
    delete x;
    return VKPBRT_OK;
}


}  // extern "C"

// ------------------------------------------------------------------------------------------------
// vkpbrt::BandedRank (include/vkpbrt/banded.hpp: the C++ layer, itself written on the C ABI above) behind the C ABI
// ------------------------------------------------------------------------------------------------
#include "../../include/vkpbrt/banded.hpp"

struct vkpbrt_banded_rank_s {
    vkpbrt::ref_ptr<vkpbrt::Context> ctx;
    vkpbrt::ref_ptr<vkpbrt::BandedRank> rank;
    int world = 1;
};

namespace {
template <class F>
int guarded(F&& f)
{
    try {
        f();
        return VKPBRT_OK;
    } catch (const std::exception& e) {
        // errors of nested C-ABI calls arrive as "vkpbrt: <last error>": keep the message, report a generic code
        return fail(VKPBRT_ERR_INVALID_ARGUMENT, e.what());
    }
}
}  // namespace

extern "C" {

int vkpbrt_stream_create(vkpbrt_context_t ctx, int high_priority, void** cuda_stream)
{
    VK_REQUIRE(ctx && cuda_stream, "vkpbrt_stream_create: null argument");
    VK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t s = nullptr;
#ifndef VKPBRT_HOSTSIM
    int lo = 0, hi = 0;
    VK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    VK_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo));
#else
    (void)high_priority;
    VK_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
#endif
    *cuda_stream = s;
    return VKPBRT_OK;
}

int vkpbrt_stream_destroy(vkpbrt_context_t ctx, void* cuda_stream)
{
    VK_REQUIRE(ctx, "null context");
    if (cuda_stream) cudaStreamDestroy((cudaStream_t)cuda_stream);
    return VKPBRT_OK;
}

int vkpbrt_banded_rank_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, int rank, int world, int use_taa, int max_disp_rows,
                              int external_inputs, void* comm_stream, vkpbrt_all_gather_fn all_gather, void* user, uint32_t timeout_ms,
                              vkpbrt_banded_rank_t* out)
{
    VK_REQUIRE(ctx && out && world >= 1 && rank >= 0 && rank < world, "vkpbrt_banded_rank_create: bad argument");
    VK_REQUIRE(world == 1 || all_gather, "vkpbrt_banded_rank_create: more than one rank needs an all_gather callback");
    VK_REQUIRE(world <= VKPBRT_HALO_MAX_PEERS, "vkpbrt_banded_rank_create: at most VKPBRT_HALO_MAX_PEERS ranks per node");
    auto* r = new vkpbrt_banded_rank_s();
    r->world = world;
    const int rc = guarded([&] {
        r->ctx = std::make_shared<vkpbrt::Context>(ctx);
        vkpbrt::BandedRank::Options opt;
        opt.use_taa = use_taa != 0;
        opt.max_disp_rows = max_disp_rows;
        opt.external_inputs = external_inputs != 0;
        opt.comm_stream = comm_stream;
        opt.timeout_ms = timeout_ms;
        vkpbrt::AllGather ag = [all_gather, user, world, rank](const std::vector<vkpbrt::PeerHandle>& mine) {
            std::vector<std::vector<vkpbrt::PeerHandle>> everyone((size_t)world, std::vector<vkpbrt::PeerHandle>(mine.size()));
            if (world == 1) { everyone[0] = mine; return everyone; }
            const uint64_t bytes = mine.size() * sizeof(vkpbrt::PeerHandle);
            std::vector<unsigned char> flat((size_t)world * bytes);
            if (all_gather(user, mine.data(), bytes, flat.data()) != 0) throw std::runtime_error("all_gather callback failed");
            for (int g = 0; g < world; ++g) std::memcpy(everyone[g].data(), flat.data() + (size_t)g * bytes, bytes);
            return everyone;
        };
        r->rank = vkpbrt::BandedRank::create(r->ctx, rank, world, (int)width, (int)height, opt, ag);
    });
    if (rc) { delete r; return rc; }
    *out = r;
    return VKPBRT_OK;
}

int vkpbrt_banded_rank_input_rows(vkpbrt_banded_rank_t r, int* row_begin, int* row_end)
{
    VK_REQUIRE(r && row_begin && row_end, "null argument");
    const vkpbrt::Rows rows = r->rank->input_rows();
    *row_begin = rows.lo; *row_end = rows.hi;
    return VKPBRT_OK;
}

int vkpbrt_banded_rank_block_rows(vkpbrt_banded_rank_t r, int* boundaries)
{
    VK_REQUIRE(r && boundaries, "null argument");
    for (int g = 0; g < r->world; ++g) boundaries[g] = r->rank->plan.block_rows(g).lo;
    boundaries[r->world] = r->rank->plan.block_rows(r->world - 1).hi;
    return VKPBRT_OK;
}

int vkpbrt_banded_rank_owned_rows(vkpbrt_banded_rank_t r, uint32_t frame, int* row_begin, int* row_end)
{
    VK_REQUIRE(r && row_begin && row_end, "null argument");
    const vkpbrt::Rows rows = r->rank->owned_rows((int)frame);
    *row_begin = rows.lo; *row_end = rows.hi;
    return VKPBRT_OK;
}

int vkpbrt_banded_rank_bind_inputs(vkpbrt_banded_rank_t r, void* depth, void* normal, void* albedo, void* illumination)
{
    VK_REQUIRE(r, "null argument");
    return guarded([&] { r->rank->bind_inputs(depth, normal, albedo, illumination); });
}

int vkpbrt_banded_rank_run_frame(vkpbrt_banded_rank_t r, uint32_t frame, const float* camera)
{
    VK_REQUIRE(r && camera, "null argument");
    return guarded([&] { r->rank->run_frame((int)frame, camera); });
}

int vkpbrt_banded_rank_flush(vkpbrt_banded_rank_t r)
{
    VK_REQUIRE(r, "null argument");
    return guarded([&] { r->rank->flush(); });
}

int vkpbrt_banded_rank_check(vkpbrt_banded_rank_t r)
{
    VK_REQUIRE(r, "null argument");
    return guarded([&] { r->rank->check_errors(); });
}

int vkpbrt_banded_rank_image(vkpbrt_banded_rank_t r, uint32_t which, vkpbrt_image_t* out)
{
    VK_REQUIRE(r && out, "null argument");
    switch (which) {
    case VKPBRT_BANDED_IMAGE_FINAL: *out = r->rank->final_image->handle; break;
    case VKPBRT_BANDED_IMAGE_DENOISER_FINAL: *out = r->rank->denoiser_final->handle; break;
    case VKPBRT_BANDED_IMAGE_DENOISED: *out = r->rank->denoised->handle; break;
    default: return fail(VKPBRT_ERR_INVALID_ARGUMENT, "vkpbrt_banded_rank_image: unknown image");
    }
    return VKPBRT_OK;
}

int vkpbrt_banded_rank_stats(vkpbrt_banded_rank_t r, uint64_t spin_ns[8], uint64_t* bytes_pushed)
{
    VK_REQUIRE(r && spin_ns && bytes_pushed, "null argument");
    return guarded([&] {
        const auto s = r->rank->spin_ns();
        for (int g = 0; g < 4; ++g) { spin_ns[2 * g] = s[g][0]; spin_ns[2 * g + 1] = s[g][1]; }
        *bytes_pushed = r->rank->bytes_exchanged();
    });
}

int vkpbrt_banded_rank_destroy(vkpbrt_banded_rank_t r)
{
    if (!r) return VKPBRT_OK;
    r->rank.reset();
    r->ctx.reset();
    delete r;
    return VKPBRT_OK;
}

}  // extern "C"
