/*
 * vkpbrt_b200.h -- C ABI of the B200-native denoising modules (libvkpbrt_b200.so).
 *
 * Drop-in boundary for VulkanPBRT's data-parallel denoising path.  The reference has no FFI
 * layer: its callers use C++ render-module classes directly.  This
 * header is what those classes bind to in the replacement; the C++ headers under include/vkpbrt/ rebuild the
 * reference's class surface (same names, constructor argument order, method names) on top of
 * it.  Every entry point cites the reference interface it replaces (paths relative to the
 * reference root).
 *
 * Model:
 *   - plain C, opaque handles, int status returns (0 == VKPBRT_OK), no exceptions cross the
 *     ABI; vkpbrt_last_error() returns the message of the calling thread's last failure.
 *   - all image memory is pitch-linear device memory (cudaMalloc, or imported Vulkan
 *     external memory, see vkpbrt_import_external_memory_fd).  TILING_OPTIMAL images are not
 *     importable as linear pointers; the Vulkan side must allocate exportable linear images
 *     or buffers.
 *   - "record" == enqueue on the context's CUDA stream, in call order.  Stream order replaces
 *     the reference's COMPUTE->COMPUTE pipeline barriers (denoisers/BMFR.cpp:203-230).  The
 *     C++ layer keeps the reference's record-once / replay-per-frame command list on top.
 *   - there is no CPU fallback: every *_record call launches sm_100a kernels or fails.
 *   - lifetimes: images are reference counted (vkpbrt_image_retain / _release); a buffer bundle (g-buffer, illumination,
 *     accumulation buffer) retains its images.  A MODULE (accumulator, bmfr, bfr, blender, taa, converter) borrows the
 *     bundles and the context it was created from: they must outlive it, and modules are destroyed before the context.
 *     The C++ classes of include/vkpbrt/ hold the references that guarantee this (ref_ptr members, and the recorded
 *     command closures keep their module alive, as the reference's command graph does).
 */
#ifndef VKPBRT_B200_H
#define VKPBRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VKPBRT_API
#else
#define VKPBRT_API __attribute__((visibility("default")))
#endif

/* ---------------------------------------------------------------------------------------- */
/* status                                                                                    */
/* ---------------------------------------------------------------------------------------- */
enum {
    VKPBRT_OK = 0,
    VKPBRT_ERR_INVALID_ARGUMENT = 1,
    VKPBRT_ERR_CUDA = 2,              /* a CUDA runtime call failed; message has cudaGetErrorString */
    VKPBRT_ERR_UNSUPPORTED = 3,       /* e.g. block size / fitting kernel combination            */
    VKPBRT_ERR_WRONG_BUFFER_TYPE = 4, /* denoisers/BMFR.cpp:17-22, BFR.cpp:15-20: reference prints and
                                         half-constructs; the replacement rejects               */
    VKPBRT_ERR_NOT_COMPILED = 5,      /* record() before compile()                               */
    VKPBRT_ERR_MISSING_MATRICES = 6,  /* Accumulator.cpp:89-94 (separate matrices requested, absent) */
    VKPBRT_ERR_NO_DEVICE = 7          /* no CUDA device / not sm_100: the library never falls back */
};

VKPBRT_API const char* vkpbrt_last_error(void);
VKPBRT_API const char* vkpbrt_version(void);

/* ---------------------------------------------------------------------------------------- */
/* context  ~ vsg::Context& handed to compile()/update_image_layouts()                       */
/* ---------------------------------------------------------------------------------------- */
typedef struct vkpbrt_context_s* vkpbrt_context_t;

/* stream: a cudaStream_t owned by the caller, or NULL to let the context create its own.     */
VKPBRT_API int vkpbrt_context_create(int device, void* cuda_stream, vkpbrt_context_t* out);
VKPBRT_API int vkpbrt_context_destroy(vkpbrt_context_t ctx);
VKPBRT_API int vkpbrt_context_synchronize(vkpbrt_context_t ctx);
VKPBRT_API int vkpbrt_context_stream(vkpbrt_context_t ctx, void** cuda_stream);
/* number of kernels this context has launched so far (bench.py's gpu_launches)               */
VKPBRT_API int vkpbrt_context_launch_count(vkpbrt_context_t ctx, uint64_t* out);
/* vsg::inverse(const mat4&) as the reference's host code uses it (Accumulator.cpp:100, the BMFR-dataset matrix import
 * RenderIO.cpp:639-664; external/vsg/src/vsg/maths/maths_transform.cpp:36-156): column-major 4x4, binary32, the same
 * operations in the same order, so a host that lets this library invert its matrices feeds the modules the bits the
 * reference would.  Pure host code: needs no device.  m and inverse may alias. */
VKPBRT_API int vkpbrt_mat4_inverse(const float m[16], float inverse[16]);
/* CUDA devices of this process and their 16-byte UUIDs: a Vulkan host picks the CUDA device whose UUID equals
 * VkPhysicalDeviceIDProperties::deviceUUID of the VkPhysicalDevice it renders on (include/vkpbrt/vk_interop.hpp);
 * external memory can only be imported on the device that exported it. */
VKPBRT_API int vkpbrt_device_count(int* count);
VKPBRT_API int vkpbrt_device_uuid(int device, uint8_t uuid[16]);
/* Device-side self check (tests): evaluates the tone-map quantiser of bmfrPost.comp:121-123 / bfr.comp:306-308 for
 * EVERY non-negative binary32 input both ways -- the kernels' threshold search and the pow form -- and returns the
 * number of inputs on which they differ (must be 0) and the smallest such bit pattern. */
VKPBRT_API int vkpbrt_debug_tonemap_sweep(vkpbrt_context_t ctx, uint64_t* mismatches, uint32_t* first_mismatch);

/* ---------------------------------------------------------------------------------------- */
/* images  ~ vsg::ref_ptr<vsg::DescriptorImage>                                              */
/* ---------------------------------------------------------------------------------------- */
typedef enum {
    VKPBRT_FORMAT_UNDEFINED = 0,
    VKPBRT_FORMAT_R32_SFLOAT = 1,          /* GBuffer.cpp:60  depth, AccumulationBuffer.cpp:277 prev_depth */
    VKPBRT_FORMAT_R32G32_SFLOAT = 2,       /* GBuffer.cpp:77  normal (theta, phi)                       */
    VKPBRT_FORMAT_R8G8B8A8_UNORM = 3,      /* GBuffer.cpp:94,111 material/albedo; Taa.cpp:24 history       */
    VKPBRT_FORMAT_B8G8R8A8_UNORM = 4,      /* BMFR.cpp:81, BFR.cpp, BFRBlender.cpp, Taa.cpp:42 finals       */
    VKPBRT_FORMAT_R16G16_SFLOAT = 5,       /* AccumulationBuffer.cpp:303 motion (normalised uv)           */
    VKPBRT_FORMAT_R8_UNORM = 6,            /* AccumulationBuffer.cpp:251,264 spp / prev_spp                */
    VKPBRT_FORMAT_R16G16B16A16_SFLOAT = 7, /* IlluminationBuffer.cpp:235, BMFR.cpp:59-64 denoised x2 layers */
    VKPBRT_FORMAT_R32G32B32A32_SFLOAT = 8, /* IlluminationBuffer.cpp:260-282 raw 1-spp illumination        */
    VKPBRT_FORMAT_R16_SFLOAT = 9,          /* BMFR.cpp:99-104 feature buffer, 13 layers                    */
} vkpbrt_format;

typedef struct vkpbrt_image_s* vkpbrt_image_t;

typedef struct {
    void* data;          /* CURRENT device pointer (ping-pong images flip it per frame)        */
    uint32_t format;     /* vkpbrt_format                                                      */
    uint32_t width, height, layers;
    uint64_t row_pitch;  /* bytes; always width * texel size (tightly packed)                  */
    uint64_t layer_pitch;
    uint64_t size_bytes; /* layers * layer_pitch                                               */
    int32_t owned;       /* 1: allocated by compile(); 0: wraps caller memory                   */
} vkpbrt_image_info;

VKPBRT_API uint32_t vkpbrt_format_texel_size(uint32_t format);
/* describes an image; memory is allocated (and zeroed) by vkpbrt_image_compile, mirroring
 * vsg::DescriptorImage::compile(context) (BMFR.cpp:172-178) */
VKPBRT_API int vkpbrt_image_create(vkpbrt_context_t ctx, uint32_t format, uint32_t width, uint32_t height,
                                   uint32_t layers, vkpbrt_image_t* out);
/* wraps caller-owned device memory (Vulkan-imported or another allocator); tightly packed */
VKPBRT_API int vkpbrt_image_wrap(vkpbrt_context_t ctx, uint32_t format, uint32_t width, uint32_t height,
                                 uint32_t layers, void* device_ptr, vkpbrt_image_t* out);
VKPBRT_API int vkpbrt_image_set_data(vkpbrt_image_t img, void* device_ptr); /* wrapped images only */
VKPBRT_API int vkpbrt_image_compile(vkpbrt_image_t img);
VKPBRT_API int vkpbrt_image_info_get(vkpbrt_image_t img, vkpbrt_image_info* out);
/* async copies on the context stream (pin the host buffer for overlap) */
VKPBRT_API int vkpbrt_image_upload(vkpbrt_image_t img, const void* host, uint64_t bytes);
VKPBRT_API int vkpbrt_image_download(vkpbrt_image_t img, void* host, uint64_t bytes);
VKPBRT_API int vkpbrt_image_clear(vkpbrt_image_t img);
/* device-to-device copy of a whole image on the context stream: BFRBlender::copy_final_image / Taa::copy_final_image
 * (denoisers/BFRBlender.cpp:91-124, Taa.cpp:108-141: vkCmdCopyImage between two images of equal extent and texel
 * size; the pipeline barriers around it are stream order here).  Formats must have the same texel size, extents and
 * layer counts must match. */
VKPBRT_API int vkpbrt_image_copy_record(vkpbrt_image_t src, vkpbrt_image_t dst);
VKPBRT_API int vkpbrt_image_retain(vkpbrt_image_t img);
VKPBRT_API int vkpbrt_image_release(vkpbrt_image_t img);

/* ---------------------------------------------------------------------------------------- */
/* buffer bundles                                                                            */
/* ---------------------------------------------------------------------------------------- */
/* GBuffer (source/buffers/GBuffer.hpp:12-21): public members depth, normal, material, albedo */
typedef struct vkpbrt_gbuffer_s* vkpbrt_gbuffer_t;
typedef enum { VKPBRT_GBUFFER_DEPTH = 0, VKPBRT_GBUFFER_NORMAL = 1, VKPBRT_GBUFFER_MATERIAL = 2, VKPBRT_GBUFFER_ALBEDO = 3 } vkpbrt_gbuffer_member;
VKPBRT_API int vkpbrt_gbuffer_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, vkpbrt_gbuffer_t* out);
VKPBRT_API int vkpbrt_gbuffer_create_from_images(vkpbrt_context_t ctx, vkpbrt_image_t depth, vkpbrt_image_t normal,
                                                 vkpbrt_image_t material, vkpbrt_image_t albedo, vkpbrt_gbuffer_t* out);
VKPBRT_API int vkpbrt_gbuffer_compile(vkpbrt_gbuffer_t g);
VKPBRT_API int vkpbrt_gbuffer_image(vkpbrt_gbuffer_t g, uint32_t member, vkpbrt_image_t* out); /* borrowed */
VKPBRT_API int vkpbrt_gbuffer_destroy(vkpbrt_gbuffer_t g);

/* IlluminationBuffer family (source/buffers/IlluminationBuffer.hpp:14-78) */
typedef struct vkpbrt_illumination_buffer_s* vkpbrt_illumination_buffer_t;
typedef enum {
    VKPBRT_ILLUMINATION_FINAL = 0,             /* IlluminationBufferFinal: rejected by the denoisers     */
    VKPBRT_ILLUMINATION_DEMODULATED = 1,       /* 2 x rgba16f: illumination, illuminationSquared          */
    VKPBRT_ILLUMINATION_DEMODULATED_FLOAT = 2, /* 1 x rgba32f: raw 1-spp demodulated illumination         */
    VKPBRT_ILLUMINATION_FINAL_DEMODULATED = 3, /* rejected by the denoisers                               */
} vkpbrt_illumination_type;
VKPBRT_API int vkpbrt_illumination_buffer_create(vkpbrt_context_t ctx, uint32_t type, uint32_t width,
                                                 uint32_t height, vkpbrt_illumination_buffer_t* out);
/* wraps caller-provided images (e.g. Vulkan-imported raw illumination) in a buffer of the given type */
VKPBRT_API int vkpbrt_illumination_buffer_create_from_images(vkpbrt_context_t ctx, uint32_t type, const vkpbrt_image_t* images,
                                                             uint32_t count, vkpbrt_illumination_buffer_t* out);
VKPBRT_API int vkpbrt_illumination_buffer_compile(vkpbrt_illumination_buffer_t b);
VKPBRT_API int vkpbrt_illumination_buffer_type(vkpbrt_illumination_buffer_t b, uint32_t* type, uint32_t* image_count);
VKPBRT_API int vkpbrt_illumination_buffer_image(vkpbrt_illumination_buffer_t b, uint32_t index, vkpbrt_image_t* out);
VKPBRT_API int vkpbrt_illumination_buffer_destroy(vkpbrt_illumination_buffer_t b);

/* AccumulationBuffer (source/buffers/AccumulationBuffer.hpp:13-24) */
typedef struct vkpbrt_accumulation_buffer_s* vkpbrt_accumulation_buffer_t;
typedef enum {
    VKPBRT_ACC_PREV_ILLU = 0, VKPBRT_ACC_PREV_ILLU_SQUARED = 1, VKPBRT_ACC_PREV_DEPTH = 2,
    VKPBRT_ACC_PREV_NORMAL = 3, VKPBRT_ACC_SPP = 4, VKPBRT_ACC_PREV_SPP = 5, VKPBRT_ACC_MOTION = 6,
    VKPBRT_ACC_NEXT_DEPTH = 7 /* not in the reference: the depth history k_accumulate is writing (becomes prev_depth at copy_to_back) */
} vkpbrt_accumulation_member;
VKPBRT_API int vkpbrt_accumulation_buffer_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height,
                                                 vkpbrt_accumulation_buffer_t* out);
VKPBRT_API int vkpbrt_accumulation_buffer_compile(vkpbrt_accumulation_buffer_t b);
VKPBRT_API int vkpbrt_accumulation_buffer_image(vkpbrt_accumulation_buffer_t b, uint32_t member, vkpbrt_image_t* out);
/* AccumulationBuffer::copy_to_back_images (AccumulationBuffer.cpp:72-244): end-of-frame history
 * rotation depth->prev_depth, spp->prev_spp, illumination[0]->prev_illu.  The five image copies of
 * the reference (58 B/pixel) become pointer swaps: spp/prev_spp and illumination/prev_illu are
 * exchanged when both sides are library-owned (device copy otherwise); prev_depth is produced by
 * the accumulate kernel itself (it already reads depth) into a ping-pong pair that is flipped
 * here.  prev_normal / prev_illu_squared are not maintained: no shader on the path reads them
 * (accumulator.comp:82-83 is commented out, :104 never writes illuminationSquared). */
VKPBRT_API int vkpbrt_accumulation_buffer_copy_to_back_images(vkpbrt_accumulation_buffer_t b, vkpbrt_gbuffer_t g,
                                                              vkpbrt_illumination_buffer_t illumination);
VKPBRT_API int vkpbrt_accumulation_buffer_destroy(vkpbrt_accumulation_buffer_t b);

/* ---------------------------------------------------------------------------------------- */
/* per-frame constants                                                                       */
/* ---------------------------------------------------------------------------------------- */
/* RayTracingPushConstants (source/renderModules/PipelineStructs.hpp:6-13), column-major mat4s */
typedef struct {
    float view_inverse[16];
    float proj_inverse[16];
    float prev_view[16];
    uint32_t frame_number;
    uint32_t sample_number;
} vkpbrt_push_constants;

/* CameraMatrices (source/io/RenderIO.hpp:28-34) */
typedef struct {
    float view[16];     /* combined view-projection when has_proj == 0 */
    float inv_view[16];
    int32_t has_proj;   /* std::optional<mat4> proj, inv_proj */
    float proj[16];
    float inv_proj[16];
} vkpbrt_camera_matrices;

/* ---------------------------------------------------------------------------------------- */
/* Accumulator  (source/renderModules/Accumulator.hpp:15-25, Accumulator.cpp:4-117;          */
/*               kernel: shaders/accumulator.comp:33-104)                                    */
/* ---------------------------------------------------------------------------------------- */
typedef struct vkpbrt_accumulator_s* vkpbrt_accumulator_t;
/* Accumulator(g_buffer, illumination_buffer, separate_matrices, work_width = 16, work_height = 16):
 * creates and owns an AccumulationBuffer and an IlluminationBufferDemodulated. */
VKPBRT_API int vkpbrt_accumulator_create(vkpbrt_context_t ctx, vkpbrt_gbuffer_t g, vkpbrt_illumination_buffer_t illumination,
                                         int separate_matrices, int work_width, int work_height,
                                         vkpbrt_accumulator_t* out);
VKPBRT_API int vkpbrt_accumulator_compile_images(vkpbrt_accumulator_t a);              /* compile_images()          */
VKPBRT_API int vkpbrt_accumulator_accumulated_illumination(vkpbrt_accumulator_t a, vkpbrt_illumination_buffer_t* out);
VKPBRT_API int vkpbrt_accumulator_accumulation_buffer(vkpbrt_accumulator_t a, vkpbrt_accumulation_buffer_t* out);
VKPBRT_API int vkpbrt_accumulator_set_camera_matrices(vkpbrt_accumulator_t a, int frame_index,
                                                      const vkpbrt_camera_matrices* cur, const vkpbrt_camera_matrices* prev);
/* add_dispatch_to_command_graph(): one launch of the accumulate kernel */
VKPBRT_API int vkpbrt_accumulator_record(vkpbrt_accumulator_t a);
/* restrict the dispatch to image rows [row_begin, row_end) (multi-GPU band sharding) */
VKPBRT_API int vkpbrt_accumulator_set_row_range(vkpbrt_accumulator_t a, int row_begin, int row_end);
/* debug / test switch: non-zero = run the one-pixel-per-thread kernel with the IEEE library routines instead of the
 * packed two-pixel kernel (both are bit-identical; the tests run both) */
VKPBRT_API int vkpbrt_accumulator_set_force_scalar(vkpbrt_accumulator_t a, int enable);
/* Band-sharded runs: a rank holds history rows only within `rows` (+ 1 bilinear row) of the rows it computes.  With the
 * guard on (rows > 0) every reprojection tap further away than that -- other than through the sampler's REPEAT wrap --
 * is counted instead of silently reading rows the rank never received; ..._displacement_violations() synchronises
 * the context's stream and returns the count since the guard was switched on. */
VKPBRT_API int vkpbrt_accumulator_set_max_displacement_rows(vkpbrt_accumulator_t a, int rows);
VKPBRT_API int vkpbrt_accumulator_displacement_violations(vkpbrt_accumulator_t a, uint32_t* count);
VKPBRT_API int vkpbrt_accumulator_destroy(vkpbrt_accumulator_t a);

/* ---------------------------------------------------------------------------------------- */
/* BMFR  (source/renderModules/denoisers/BMFR.hpp:17-25, BMFR.cpp:5-234;                     */
/*        kernels: shaders/bmfrPre.comp, bmfrFit.comp, bmfrPost.comp)                        */
/* ---------------------------------------------------------------------------------------- */
typedef struct vkpbrt_bmfr_s* vkpbrt_bmfr_t;
/* BMFR(width, height, work_width, work_height, g_buffer, illu_buffer, acc_buffer, fitting_kernel = 256).
 * Supported (work, fitting_kernel): (32,256) (16,256) (8,64) -- the combinations
 * util/DenoiserUtils.cpp:78-124 instantiates. */
VKPBRT_API int vkpbrt_bmfr_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, uint32_t work_width,
                                  uint32_t work_height, vkpbrt_gbuffer_t g, vkpbrt_illumination_buffer_t illumination,
                                  vkpbrt_accumulation_buffer_t acc, uint32_t fitting_kernel, vkpbrt_bmfr_t* out);
/* the fused kernel keeps the feature matrix and the weights on chip; enable to also materialise
 * the reference's featureBuffer (r16f x13, padded) and weights (r32f x30) images for parity tests.
 * enable bit 0: those images; bit 1: every block takes the out-of-line IEEE-division fit (the path a block
 * falls back to when an operand leaves the range the reciprocal division is proven exact for) */
VKPBRT_API int vkpbrt_bmfr_set_debug_outputs(vkpbrt_bmfr_t b, int enable);
/* The shaders' POSITION_TYPE specialisation constant (bmfrGeneral.comp:30-31; bmfrPre.comp:36-76, bmfrPost.comp:31-71):
 * 0 POSITION_DEPTH (default -- the only mode the reference's BMFR.cpp instantiates), 1 POSITION_WORLD_DEPTH_NORM,
 * 2 POSITION_WORLD.  The WORLD modes build the three position features from the camera ray through the pixel and read
 * view_inverse / proj_inverse of the push constants handed to vkpbrt_bmfr_record. */
#define VKPBRT_BMFR_POSITION_DEPTH 0
#define VKPBRT_BMFR_POSITION_WORLD_DEPTH_NORM 1
#define VKPBRT_BMFR_POSITION_WORLD 2
VKPBRT_API int vkpbrt_bmfr_set_position_type(vkpbrt_bmfr_t b, int position_type);
/* Side lanes.  Denoisers of one frame that do not depend on each other (the three block sizes of X8X16X32) can run
 * concurrently: lane 0 (default) records on the context's stream, lanes 1 and 2 on streams of their own that are forked
 * from the context's stream at record() and joined back into it before whatever is recorded on it next (blender, TAA,
 * copies, synchronize).  The reference serialises them with COMPUTE->COMPUTE barriers (DenoiserUtils.cpp:48-70). */
VKPBRT_API int vkpbrt_bmfr_set_lane(vkpbrt_bmfr_t b, int lane);
VKPBRT_API int vkpbrt_bmfr_compile(vkpbrt_bmfr_t b);
VKPBRT_API int vkpbrt_bmfr_record(vkpbrt_bmfr_t b, const vkpbrt_push_constants* pc); /* pre+fit+post, one launch */
VKPBRT_API int vkpbrt_bmfr_set_block_row_range(vkpbrt_bmfr_t b, int block_row_begin, int block_row_end);
VKPBRT_API int vkpbrt_bmfr_final_image(vkpbrt_bmfr_t b, vkpbrt_image_t* out);        /* get_final_descriptor_image() */
typedef enum { VKPBRT_BMFR_IMAGE_DENOISED = 0, VKPBRT_BMFR_IMAGE_FEATURES = 1, VKPBRT_BMFR_IMAGE_WEIGHTS = 2 } vkpbrt_bmfr_image;
VKPBRT_API int vkpbrt_bmfr_image_get(vkpbrt_bmfr_t b, uint32_t which, vkpbrt_image_t* out);
VKPBRT_API int vkpbrt_bmfr_destroy(vkpbrt_bmfr_t b);

/* ---------------------------------------------------------------------------------------- */
/* BFR  (denoisers/BFR.hpp:11-18, BFR.cpp:6-142; kernel: shaders/bfr.comp:202-309)           */
/* ---------------------------------------------------------------------------------------- */
typedef struct vkpbrt_bfr_s* vkpbrt_bfr_t;
VKPBRT_API int vkpbrt_bfr_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, uint32_t work_width,
                                 uint32_t work_height, vkpbrt_gbuffer_t g, vkpbrt_illumination_buffer_t illumination,
                                 vkpbrt_accumulation_buffer_t acc, vkpbrt_bfr_t* out);
VKPBRT_API int vkpbrt_bfr_set_lane(vkpbrt_bfr_t b, int lane);     /* see vkpbrt_bmfr_set_lane */
VKPBRT_API int vkpbrt_bfr_compile(vkpbrt_bfr_t b);
VKPBRT_API int vkpbrt_bfr_record(vkpbrt_bfr_t b, const vkpbrt_push_constants* pc);
VKPBRT_API int vkpbrt_bfr_final_image(vkpbrt_bfr_t b, vkpbrt_image_t* out);
VKPBRT_API int vkpbrt_bfr_denoised_image(vkpbrt_bfr_t b, vkpbrt_image_t* out);
VKPBRT_API int vkpbrt_bfr_destroy(vkpbrt_bfr_t b);

/* ---------------------------------------------------------------------------------------- */
/* BFRBlender  (denoisers/BFRBlender.hpp:9-18, BFRBlender.cpp:5-130;                         */
/*              kernel: shaders/bfrBlender.comp:22-68)                                       */
/* ---------------------------------------------------------------------------------------- */
typedef struct vkpbrt_bfr_blender_s* vkpbrt_bfr_blender_t;
/* BFRBlender(width, height, average_image, average_squared_image, denoised0, denoised1, denoised2,
 *            work_width = 16, work_height = 16, filter_radius = 2) */
VKPBRT_API int vkpbrt_bfr_blender_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, vkpbrt_image_t average,
                                         vkpbrt_image_t average_squared, vkpbrt_image_t denoised0,
                                         vkpbrt_image_t denoised1, vkpbrt_image_t denoised2, uint32_t work_width,
                                         uint32_t work_height, uint32_t filter_radius, vkpbrt_bfr_blender_t* out);
VKPBRT_API int vkpbrt_bfr_blender_compile(vkpbrt_bfr_blender_t b);
VKPBRT_API int vkpbrt_bfr_blender_record(vkpbrt_bfr_blender_t b);
VKPBRT_API int vkpbrt_bfr_blender_final_image(vkpbrt_bfr_blender_t b, vkpbrt_image_t* out);
VKPBRT_API int vkpbrt_bfr_blender_destroy(vkpbrt_bfr_blender_t b);

/* ---------------------------------------------------------------------------------------- */
/* Taa  (source/renderModules/Taa.hpp:14-20, Taa.cpp:4-147; kernel: shaders/taa.comp:48-104) */
/* ---------------------------------------------------------------------------------------- */
typedef struct vkpbrt_taa_s* vkpbrt_taa_t;
/* Taa(width, height, work_width, work_height, g_buffer, acc_buffer, denoised) */
VKPBRT_API int vkpbrt_taa_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, uint32_t work_width,
                                 uint32_t work_height, vkpbrt_gbuffer_t g, vkpbrt_accumulation_buffer_t acc,
                                 vkpbrt_image_t denoised, vkpbrt_taa_t* out);
/* 0 (default): reproduce the reference's R/B-swapped history; 1: fix it */
VKPBRT_API int vkpbrt_taa_set_fix_swizzle(vkpbrt_taa_t t, int fix);
/* debug / test switch: the one-pixel-per-thread kernel (bit-identical to the default two-column kernel) */
VKPBRT_API int vkpbrt_taa_set_force_scalar(vkpbrt_taa_t t, int enable);
/* debug / test switch: rows a warp of the two-column kernel walks per launch (0 = automatic: 16, less when the launch would
 * otherwise leave most of the GPU's warp slots empty).  Results do not depend on it. */
VKPBRT_API int vkpbrt_taa_set_strip_rows(vkpbrt_taa_t t, int rows);
VKPBRT_API int vkpbrt_taa_compile(vkpbrt_taa_t t);
/* dispatch + final->history hand-over (Taa.cpp:99-107).  The reference never pushes constants for
 * TAA and inherits the denoiser's (App. C-10); here they are passed explicitly. */
VKPBRT_API int vkpbrt_taa_record(vkpbrt_taa_t t, const vkpbrt_push_constants* pc);
VKPBRT_API int vkpbrt_taa_set_row_range(vkpbrt_taa_t t, int row_begin, int row_end);
/* One frame's TAA in several launches over disjoint row ranges (band-sharded runs: the rows that do not need the
 * neighbour's halo row run while that row is still in flight).  Every part reads the same history; the final -> history
 * hand-over (Taa.cpp:106) happens with the part that passes last != 0. */
VKPBRT_API int vkpbrt_taa_record_part(vkpbrt_taa_t t, const vkpbrt_push_constants* pc, int row_begin, int row_end, int last);
/* the same with two disjoint row ranges in ONE launch (either may be empty): the first and the last row of a band */
VKPBRT_API int vkpbrt_taa_record_parts(vkpbrt_taa_t t, const vkpbrt_push_constants* pc, int row_begin, int row_end, int row_begin2,
                                       int row_end2, int last);
VKPBRT_API int vkpbrt_taa_final_image(vkpbrt_taa_t t, vkpbrt_image_t* out);
VKPBRT_API int vkpbrt_taa_history_image(vkpbrt_taa_t t, vkpbrt_image_t* out);
VKPBRT_API int vkpbrt_taa_destroy(vkpbrt_taa_t t);

/* ---------------------------------------------------------------------------------------- */
/* FormatConverter  (source/renderModules/FormatConverter.hpp:4-21, FormatConverter.cpp:4-93;  */
/*                   kernel: shaders/formatConverter.comp:1-13)                                 */
/* The step VulkanPBRT.cpp:476-484 appends when the final image is not B8G8R8A8_UNORM.          */
/* ---------------------------------------------------------------------------------------- */
typedef struct vkpbrt_format_converter_s* vkpbrt_format_converter_t;
/* FormatConverter(src_image, dst_format, work_width = 16, work_height = 16); dst_format must be
 * VKPBRT_FORMAT_B8G8R8A8_UNORM ("FormatConverter::Unknown format" otherwise, as in the reference) */
VKPBRT_API int vkpbrt_format_converter_create(vkpbrt_context_t ctx, vkpbrt_image_t src_image, uint32_t dst_format,
                                              uint32_t work_width, uint32_t work_height, vkpbrt_format_converter_t* out);
VKPBRT_API int vkpbrt_format_converter_compile_images(vkpbrt_format_converter_t f);
VKPBRT_API int vkpbrt_format_converter_record(vkpbrt_format_converter_t f);       /* add_dispatch_to_command_graph */
VKPBRT_API int vkpbrt_format_converter_final_image(vkpbrt_format_converter_t f, vkpbrt_image_t* out);   /* final_image */
VKPBRT_API int vkpbrt_format_converter_destroy(vkpbrt_format_converter_t f);

/* Offline sequences (source/io/RenderIO.cpp:71-158): the conversions GBufferIO::import_g_buffer_position runs on the host
 * between what the files hold and what the G-buffer holds, as ONE launch on the device: world position -> Euclidean
 * distance to the eye (:101-120; eye = column 2 of inv_view divided by its w, the convention of the offline, combined
 * matrices), cartesian normal -> (acos(n.z), atan2(n.y, n.x)) (:160-178), float albedo * 255 -> rgba8, truncating
 * (:180-195).  The planes are compiled rgba32f images of the g-buffer's size; any of them may be NULL (left alone). */
VKPBRT_API int vkpbrt_gbuffer_import_record(vkpbrt_gbuffer_t g, vkpbrt_image_t position, const float* inv_view,
                                            vkpbrt_image_t normal, vkpbrt_image_t albedo);

/* Producer-side convention (shaders/ptRaygen.rgen:81-88, DEMOD_ILLUMINATION_FLOAT): what a CUDA / OptiX path tracer has
 * to hand to the Accumulator as IlluminationBufferDemodulatedFloat.  demodulated = min(clamp(radiance, 0, 10) /
 * (albedo + 1e-6), 1e3) where the primary ray hit something, clamp(radiance, 0, 10) where position_x is infinite. */
VKPBRT_API int vkpbrt_demodulate_record(vkpbrt_context_t ctx, vkpbrt_image_t radiance, vkpbrt_image_t albedo,
                                        vkpbrt_image_t position_x, vkpbrt_image_t demodulated);

/* ---------------------------------------------------------------------------------------- */
/* Vulkan interop (north star: VK_KHR_external_memory_fd / VK_KHR_external_semaphore_fd)      */
/* ---------------------------------------------------------------------------------------- */
typedef struct vkpbrt_external_memory_s* vkpbrt_external_memory_t;
typedef struct vkpbrt_external_semaphore_s* vkpbrt_external_semaphore_t;
/* imports an opaque fd exported from a VkDeviceMemory (ownership of fd passes to CUDA) and maps
 * [offset, offset+size) as a linear device pointer usable with vkpbrt_image_wrap */
VKPBRT_API int vkpbrt_import_external_memory_fd(vkpbrt_context_t ctx, int fd, uint64_t allocation_size, uint64_t offset,
                                                uint64_t size, vkpbrt_external_memory_t* out, void** device_ptr);
/* as above; dedicated != 0 when the VkDeviceMemory is a dedicated allocation (VkMemoryDedicatedAllocateInfo:
 * cudaExternalMemoryDedicated has to be passed for those, and some drivers export buffers only that way) */
VKPBRT_API int vkpbrt_import_external_memory_fd_ex(vkpbrt_context_t ctx, int fd, uint64_t allocation_size, uint64_t offset,
                                                   uint64_t size, int dedicated, vkpbrt_external_memory_t* out, void** device_ptr);
VKPBRT_API int vkpbrt_external_memory_destroy(vkpbrt_external_memory_t m);
VKPBRT_API int vkpbrt_import_external_semaphore_fd(vkpbrt_context_t ctx, int fd, int timeline, vkpbrt_external_semaphore_t* out);
VKPBRT_API int vkpbrt_external_semaphore_wait(vkpbrt_external_semaphore_t s, uint64_t value);   /* on ctx stream */
VKPBRT_API int vkpbrt_external_semaphore_signal(vkpbrt_external_semaphore_t s, uint64_t value); /* on ctx stream */
VKPBRT_API int vkpbrt_external_semaphore_destroy(vkpbrt_external_semaphore_t s);

/* ---------------------------------------------------------------------------------------- */
/* Band-sharded multi-GPU runs: halo rows over NVLink peer memory.    */
/* No reference counterpart (the reference records one device's command graph,                */
/* VulkanPBRT.cpp:551-618); the exchanged rows follow the shaders' read footprints            */
/* (accumulator.comp:75-98, bmfrPost.comp:103-118, taa.comp:44-60).  One process per GPU:      */
/* a rank exports the allocations its neighbours write into, the neighbours open them, and    */
/* each exchange point is ONE kernel that stores the rows into the receivers' HBM and then    */
/* publishes a frame counter to their flag words; receivers wait on their own flag words.     */
/* ---------------------------------------------------------------------------------------- */
#define VKPBRT_PEER_HANDLE_BYTES 64
#define VKPBRT_HALO_MAX_PEERS 8
typedef struct vkpbrt_halo_copy {       /* rows x row_bytes, pitched on both sides */
    const void* src;                    /* local device address */
    void* dst;                          /* address inside a vkpbrt_peer_open mapping (or local) */
    uint64_t src_pitch, dst_pitch;
    uint32_t row_bytes, rows;
} vkpbrt_halo_copy;
/* handle of the cudaMalloc allocation holding device_ptr + device_ptr's offset inside it (cudaIpcGetMemHandle) */
VKPBRT_API int vkpbrt_peer_export(vkpbrt_context_t ctx, const void* device_ptr, uint8_t handle[VKPBRT_PEER_HANDLE_BYTES],
                                  uint64_t* offset);
/* maps another process's allocation (cudaIpcOpenMemHandle, peer access enabled lazily); open each handle once */
VKPBRT_API int vkpbrt_peer_open(vkpbrt_context_t ctx, const uint8_t handle[VKPBRT_PEER_HANDLE_BYTES], void** base);
VKPBRT_API int vkpbrt_peer_close(vkpbrt_context_t ctx, void* base);
/* An exchange point: everything one rank does at one point of the frame for one (jitter phase, ping-pong parity),
 * built once and replayed.  `start` is ONE kernel launch on the communication stream, ordered after the work already
 * enqueued on `after_stream`:
 *   1. store `value` to every announce_flags[i]   (senders' words, peer mapped: "my halo rows may be overwritten")
 *   2. spin until every ready_flags[i] >= value    (local words the receivers announce to; n_ready may be 0: no gate)
 *   3. copy the blocks of rows of `copies`         (16-byte stores into the receivers' HBM)
 *   4. store `value` to every done_flags[i]        (receivers' words, peer mapped) once every copy has landed
 * `wait` makes a stream spin until every wait_flags[i] (local words) >= value, in front of the consuming kernel.
 * Spins are bounded by timeout_ms: on expiry an error word is set (see `stats`) and the kernel proceeds -- it never
 * hangs the GPU.  `value` must grow from call to call (a frame or sequence counter). */
typedef struct vkpbrt_halo_exchange_s* vkpbrt_halo_exchange_t;
typedef struct vkpbrt_halo_exchange_desc {
    const vkpbrt_halo_copy* copies;      /* host array; copied to the device by create */
    uint32_t n_copies;
    uint32_t* const* announce_flags;
    uint32_t n_announce;
    const uint32_t* const* ready_flags;
    uint32_t n_ready;
    uint32_t* const* done_flags;
    uint32_t n_done;
    const uint32_t* const* wait_flags;
    uint32_t n_wait;
} vkpbrt_halo_exchange_desc;
VKPBRT_API int vkpbrt_halo_exchange_create(vkpbrt_context_t ctx, const vkpbrt_halo_exchange_desc* desc, uint32_t timeout_ms,
                                           vkpbrt_halo_exchange_t* out);
VKPBRT_API int vkpbrt_halo_exchange_start(vkpbrt_halo_exchange_t x, void* comm_stream, void* after_stream, uint32_t value);
/* as vkpbrt_halo_exchange_start, but the push's gate (its ready_flags, local words) opens at gate_value instead of value:
 * "do not overwrite the receivers' rows before THEIR exchange number gate_value has reached me" -- orders a push after
 * work of the receivers that a later exchange of theirs follows, without a rendezvous on the main stream */
VKPBRT_API int vkpbrt_halo_exchange_start_gated(vkpbrt_halo_exchange_t x, void* comm_stream, void* after_stream, uint32_t value,
                                                uint32_t gate_value);
VKPBRT_API int vkpbrt_halo_exchange_wait(vkpbrt_halo_exchange_t x, void* stream, uint32_t value);
/* synchronises the device; nanoseconds spent spinning in step 2 and in `wait` since creation, and the error word */
VKPBRT_API int vkpbrt_halo_exchange_stats(vkpbrt_halo_exchange_t x, uint64_t* gate_ns, uint64_t* wait_ns, uint32_t* error);
VKPBRT_API int vkpbrt_halo_exchange_destroy(vkpbrt_halo_exchange_t x);

/* ---------------------------------------------------------------------------------------- */
/* One rank of the band-sharded chain as ONE object: the native (C++) host of the multi-GPU    */
/* path, include/vkpbrt/banded.hpp vkpbrt::BandedRank behind the C ABI.  accumulate -> BMFR b=32 */
/* [-> TAA] on this rank's band of block rows, the three halo exchange points per frame pushed  */
/* over NVLink peer memory.  A frame is one call.                                               */
/* ---------------------------------------------------------------------------------------- */
typedef struct vkpbrt_banded_rank_s* vkpbrt_banded_rank_t;
/* collective the library calls when an exchange point is first built (every rank, same order): copy `bytes` bytes of
 * `mine` to slot `rank` of `everyone` (world * bytes) on EVERY rank.  Return 0 on success. */
typedef int (*vkpbrt_all_gather_fn)(void* user, const void* mine, uint64_t bytes, void* everyone);
/* a CUDA stream for the exchange kernels (high priority); destroy with vkpbrt_stream_destroy */
VKPBRT_API int vkpbrt_stream_create(vkpbrt_context_t ctx, int high_priority, void** cuda_stream);
VKPBRT_API int vkpbrt_stream_destroy(vkpbrt_context_t ctx, void* cuda_stream);
/* external_inputs != 0: the producer's planes are bound per frame with vkpbrt_banded_rank_bind_inputs.
 * comm_stream: from vkpbrt_stream_create (NULL: the context's stream).  max_disp_rows: reprojection displacement the
 * halo covers; taps beyond it make vkpbrt_banded_rank_check fail instead of reading rows this rank never received. */
VKPBRT_API int vkpbrt_banded_rank_create(vkpbrt_context_t ctx, uint32_t width, uint32_t height, int rank, int world, int use_taa,
                                         int max_disp_rows, int external_inputs, void* comm_stream, vkpbrt_all_gather_fn all_gather,
                                         void* user, uint32_t timeout_ms, vkpbrt_banded_rank_t* out);
/* image rows of the producer's planes this rank ever touches, and the band boundaries in block rows (world + 1 ints) */
VKPBRT_API int vkpbrt_banded_rank_input_rows(vkpbrt_banded_rank_t r, int* row_begin, int* row_end);
VKPBRT_API int vkpbrt_banded_rank_block_rows(vkpbrt_banded_rank_t r, int* boundaries);
/* rows of `frame` this rank's BMFR blocks write (its part of the result) */
VKPBRT_API int vkpbrt_banded_rank_owned_rows(vkpbrt_banded_rank_t r, uint32_t frame, int* row_begin, int* row_end);
/* this frame's planes as FULL-FRAME base pointers: a band-local buffer holding rows [lo, hi) of input_rows is passed as
 * buffer - lo * row_pitch (depth r32f, normal rg32f, albedo rgba8, illumination rgba32f) */
VKPBRT_API int vkpbrt_banded_rank_bind_inputs(vkpbrt_banded_rank_t r, void* depth, void* normal, void* albedo, void* illumination);
/* camera: view, inv_view, proj, inv_proj of this frame, 4 x 16 floats, column-major */
VKPBRT_API int vkpbrt_banded_rank_run_frame(vkpbrt_banded_rank_t r, uint32_t frame, const float* camera);
VKPBRT_API int vkpbrt_banded_rank_flush(vkpbrt_banded_rank_t r);      /* stream-side wait for the halos in flight */
VKPBRT_API int vkpbrt_banded_rank_check(vkpbrt_banded_rank_t r);      /* synchronises; fails on a flag timeout / a displacement violation */
typedef enum { VKPBRT_BANDED_IMAGE_FINAL = 0, VKPBRT_BANDED_IMAGE_DENOISER_FINAL = 1, VKPBRT_BANDED_IMAGE_DENOISED = 2 } vkpbrt_banded_image;
VKPBRT_API int vkpbrt_banded_rank_image(vkpbrt_banded_rank_t r, uint32_t which, vkpbrt_image_t* out);   /* borrowed */
/* synchronises; spin_ns[group A,B,C,D][gate of the push, wait before the consumer], bytes pushed so far */
VKPBRT_API int vkpbrt_banded_rank_stats(vkpbrt_banded_rank_t r, uint64_t spin_ns[8], uint64_t* bytes_pushed);
VKPBRT_API int vkpbrt_banded_rank_destroy(vkpbrt_banded_rank_t r);

#ifdef __cplusplus
}
#endif
#endif /* VKPBRT_B200_H */
