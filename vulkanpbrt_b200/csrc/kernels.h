// kernels.h -- parameter blocks and launchers of the sm_100a kernels (host-visible part).
#pragma once
#ifdef VKPBRT_HOSTSIM
#include "hostsim.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace vkpbrt {

// ---- k_accumulate : shaders/accumulator.comp:33-104 -------------------------------------
struct AccumulateParams {
    int W, H;
    int row_begin, row_end;   // image rows processed (band sharding); full frame = [0, H)
    int separate_matrices;
    int src_is_f16;           // srcImage rgba16f (IlluminationBufferDemodulated) or rgba32f (...Float)
    uint32_t frame;
    // uniform matrices, column-major.  m_dir: "view" push constant (inverse projection in separate
    // mode, unused otherwise); inv_view; m_prev: proj*prevView (separate) or prevView (combined VP)
    float m_dir[16];
    float inv_view[16];
    float m_prev[16];
    float prev_origin[4];
    // uniform sub-expressions of the shader, evaluated once on the host with the same IEEE operations
    float uv_scale[2];        // size / (size - 0.5)                         (accumulator.comp:70)
    float rcp_size[2];        // RN(1 / size): exact (gid + .5) / size through div_by_rcp
    float cur_origin[4];      // inverseView[2] / inverseView[2].w, combined-matrix mode (:56-57)
    const void* src;              // raw 1-spp illumination
    const float* depth;           // r32f
    const float* prev_depth;      // r32f   (bilinear)
    const uint2* prev_illum;      // rgba16f (bilinear)
    const uint8_t* prev_spp;      // r8     (bilinear)
    uint32_t* motion;             // rg16f  (packed half2)
    uint8_t* spp;                 // r8
    uint2* illum;                 // rgba16f
    float* depth_history;         // r32f: next frame's prev_depth (fused copy_to_back), may be null
    float one, neg_one;           // 1.0f / -1.0f as run-time values (common.cuh: packed pairs)
    int force_scalar;             // debug: one pixel per thread with the IEEE library routines (k_accumulate_scalar)
    // band-sharded runs: a rank only holds history rows within max_disp_rows (+1 bilinear row) of the rows it computes.
    // A reprojection tap further away (other than through the REPEAT wrap at the image edge) would read rows this rank
    // never received: it is counted in *disp_violations instead of passing silently (0 / null = no check).
    int max_disp_rows;
    uint32_t* disp_violations;
};
cudaError_t launch_accumulate(const AccumulateParams& p, cudaStream_t stream);

// ---- k_bmfr_block : bmfrPre.comp + bmfrFit.comp + bmfrPost.comp, one launch ---------------
// a CUtensorMap (TMA descriptor) as plain bytes, so that this header does not need cuda.h
struct alignas(64) TmaDesc {
    unsigned char bytes[128];
};
struct BmfrParams {
    int W, H;
    int block;                    // work_width == work_height: 8, 16 or 32
    int fitting_kernel;           // threads per block row-group: 64 (b=8) or 256
    int blocks_x, blocks_y;       // W/b+2, H/b+2
    int block_row_begin, block_row_end;  // block rows processed (band sharding)
    uint32_t frame;
    int off_x, off_y;             // ivec2(vec2(b, b) * pixelOffsets[frame % 16]) (bmfrPre.comp:16), evaluated on the host
    const float* depth;
    const float2* normal;
    const uchar4* albedo;
    const uint32_t* motion;
    const uint8_t* spp;
    const uint2* noisy;           // accumulated illumination rgba16f
    const uint2* denoised_prev;   // layer (frame & 1)
    uint2* denoised_next;         // layer (frame & 1) ^ 1
    uint32_t* final_bgra;
    uint16_t* dbg_features;       // optional r16f [13][Hp][Wp]
    float* dbg_weights;           // optional r32f [30][blocks_y][blocks_x]
    int force_generic;            // debug: every block takes the out-of-line IEEE-division fit (test coverage of the cold path)
    const float* table;           // this frame's block-invariant table (bmfr.cu), 10 * block^2 floats
    float* table_next;            // where the launch's spare CTAs write the table of frame + 1 (may be null)
    float one, neg_one;           // 1.0f / -1.0f as run-time values (common.cuh: packed pairs)
    // TMA descriptors of the stage-1 input planes (32 x 32 texel boxes): blocks whose footprint lies inside the image
    // fetch their depth / normal / noisy tiles with three cp.async.bulk.tensor.2d instead of per-thread loads
    int use_tma;
    int tma_row0;                 // image row of the descriptors' row 0
    TmaDesc tma_depth, tma_normal, tma_noisy;
    // bmfrGeneral.comp:30-31 POSITION_TYPE: 0 POSITION_DEPTH (what the reference's host code runs), 1 POSITION_WORLD_DEPTH_NORM,
    // 2 POSITION_WORLD; the WORLD modes read the push constants' camera matrices (column-major)
    int position_type;
    float inv_view[16], inv_proj[16];
};
// encodes the three descriptors for the planes currently in `p` (host; needs the driver's cuTensorMapEncodeTiled).
// Leaves use_tma = 0 when a plane cannot be described (row pitch not a multiple of 16 bytes, unaligned base, block != 32).
void bmfr_encode_tma(BmfrParams& p);
constexpr size_t bmfr_table_floats(int block) { return (size_t)10 * block * block; }
cudaError_t launch_bmfr(const BmfrParams& p, cudaStream_t stream);
cudaError_t launch_bmfr_table(int block, float* table, uint32_t frame, cudaStream_t stream);

// ---- k_bfr_block : shaders/bfr.comp ------------------------------------------------------
struct BfrParams {
    int W, H;
    int block;                    // 8, 16, 32
    int blocks_x, blocks_y;
    uint32_t frame;
    // bfr.comp:134 step size factors for t = 1..40, evaluated once on the host:
    float lr_exp[40];             // exp(-K * t)
    float lr_sqrt[40];            // sqrt(1 - pow(BETA2, t))
    float lr_den[40];             // 1 - pow(BETA1, t)
    const float* depth;
    const float2* normal;
    const uchar4* albedo;
    const uint32_t* motion;
    const uint8_t* spp;
    const uint2* noisy;
    const uint2* denoised_prev;
    uint2* denoised_next;
    uint32_t* final_bgra;
};
cudaError_t launch_bfr(const BfrParams& p, cudaStream_t stream);

// ---- k_bfr_blend : shaders/bfrBlender.comp -----------------------------------------------
struct BlendParams {
    int W, H, radius;
    const uint2* average;         // rgba16f
    const uint2* average_squared; // rgba16f
    const uint32_t* denoised0;    // BGRA8 (b = 8)
    const uint32_t* denoised1;    // BGRA8 (b = 16)
    const uint32_t* denoised2;    // BGRA8 (b = 32)
    uint32_t* final_bgra;
    float one, neg_one;           // 1.0f / -1.0f as run-time values (common.cuh: packed pairs)
};
cudaError_t launch_bfr_blend(const BlendParams& p, cudaStream_t stream);

// ---- k_taa : shaders/taa.comp + Taa.cpp:106 hand-over -------------------------------------
struct TaaParams {
    int W, H;
    int row_begin, row_end;
    uint32_t frame;
    int fix_swizzle;
    const uint32_t* motion;
    const uint32_t* denoised;     // BGRA8 final of the denoiser
    const uint32_t* history;      // previous TAA final, bytes as written (BGRA8) viewed as RGBA8
    uint32_t* final_bgra;         // this frame's TAA final == next frame's history (ping-pong)
    float one, neg_one;           // 1.0f / -1.0f as run-time values (common.cuh: packed pairs)
    int force_scalar;             // debug: the one-pixel-per-thread kernel
    int row_begin2, row_end2;     // optional second row range of the same launch (empty when row_end2 <= row_begin2)
    int rows_per_warp;            // strip height of the pair kernel; <= 0: launch_taa chooses it from the size of the launch
    int strips;                   // set by launch_taa: number of strips of the first range
};
cudaError_t launch_taa(const TaaParams& p, cudaStream_t stream);

// ---- k_halo_push / k_halo_wait : band-sharded runs, halo rows over NVLink peer memory ------
struct HaloCopy {                 // rows x row_bytes, pitched on both sides (same layout as vkpbrt_halo_copy)
    const uint8_t* src;           // local
    uint8_t* dst;                 // the receiver's buffer through its peer mapping
    uint64_t src_pitch, dst_pitch;
    uint32_t row_bytes, rows;
};
constexpr int kHaloMaxPeers = 8;
struct HaloPushParams {
    const HaloCopy* copies;       // device table
    int n_copies;
    int n_announce, n_ready, n_done;
    uint32_t* announce_flags[kHaloMaxPeers];      // senders' words (peer mapped): set first
    const uint32_t* ready_flags[kHaloMaxPeers];   // local words the receivers set when their halo rows may be overwritten
    uint32_t* done_flags[kHaloMaxPeers];          // receivers' words (peer mapped), set to `value` once every copy has landed
    uint32_t value;
    uint32_t gate_value;          // what the ready flags must have reached before any copy starts
    uint32_t* counter;            // last-CTA detection, left at 0
    uint32_t* error;              // set to 1 when a spin timed out
    unsigned long long* gate_ns;  // statistics: time CTA 0 spent in the gate
    unsigned long long timeout_ns;
};
struct HaloWaitParams {
    int n;
    const uint32_t* flags[kHaloMaxPeers];
    uint32_t value;
    uint32_t* error;
    unsigned long long* wait_ns;  // statistics
    unsigned long long timeout_ns;
};
// ---- convert.cu : shaders/formatConverter.comp, ptRaygen.rgen:81-88 -----------------------------------
struct FormatConvertParams {
    int W, H;
    int src_format;               // 0 rgba32f, 1 rgba16f, 2 rgba8 unorm
    const void* src;
    uint32_t* dst_bgra;           // B8G8R8A8_UNORM
};
cudaError_t launch_format_convert(const FormatConvertParams& p, cudaStream_t stream);
struct DemodulateParams {
    int W, H;
    const float4* radiance;       // rgba32f: the path tracer's radiance estimate (finalColor before the clamp)
    const float4* albedo;         // rgba32f: diffuse + specular colour of the primary hit (curAlbedo)
    const float* position_x;      // r32f: x of the primary hit position, +-inf for a miss (rayPayload.position.x)
    float4* out;                  // rgba32f demodulated illumination (IlluminationBufferDemodulatedFloat)
};
cudaError_t launch_demodulate(const DemodulateParams& p, cudaStream_t stream);

// offline-sequence import conversions (source/io/RenderIO.cpp:101-120, :160-195); any of the three inputs may be null
struct GBufferImportParams {
    int W, H;
    float camera[3];              // eye point: column 2 of the (combined) inverse view matrix divided by its w (:109-110)
    const float4* position;       // rgba32f world position  -> depth
    const float4* normal;         // rgba32f cartesian normal -> (theta, phi)
    const float4* albedo;         // rgba32f albedo           -> rgba8 (truncating)
    float* depth;
    float2* normal_out;
    uint32_t* albedo_out;
};
cudaError_t launch_gbuffer_import(const GBufferImportParams& p, cudaStream_t stream);

// ---- device-side self checks (debug.cu) ----------------------------------------------------------
cudaError_t launch_tonemap_sweep(unsigned long long* bad, uint32_t* first_bad, cudaStream_t stream);

cudaError_t launch_halo_push(const HaloPushParams& p, int parts, cudaStream_t stream);
cudaError_t launch_halo_wait(const HaloWaitParams& p, cudaStream_t stream);

}  // namespace vkpbrt
