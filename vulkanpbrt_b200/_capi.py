"""ctypes binding of the C ABI in include/vkpbrt_b200.h (libvkpbrt_b200.so).

This is the only way the Python host layer reaches the kernels.  There is no fallback: if the
shared library is missing the import raises, and on a machine without an sm_100 GPU
``vkpbrt_context_create`` fails with VKPBRT_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libvkpbrt_b200.so"


class VkpbrtError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[vkpbrt error {code}] {message}")
        self.code = code


OK = 0
ERR_INVALID_ARGUMENT = 1
ERR_CUDA = 2
ERR_UNSUPPORTED = 3
ERR_WRONG_BUFFER_TYPE = 4
ERR_NOT_COMPILED = 5
ERR_MISSING_MATRICES = 6
ERR_NO_DEVICE = 7

FORMAT_R32_SFLOAT = 1
FORMAT_R32G32_SFLOAT = 2
FORMAT_R8G8B8A8_UNORM = 3
FORMAT_B8G8R8A8_UNORM = 4
FORMAT_R16G16_SFLOAT = 5
FORMAT_R8_UNORM = 6
FORMAT_R16G16B16A16_SFLOAT = 7
FORMAT_R32G32B32A32_SFLOAT = 8
FORMAT_R16_SFLOAT = 9

GBUFFER_DEPTH, GBUFFER_NORMAL, GBUFFER_MATERIAL, GBUFFER_ALBEDO = range(4)
ILLUMINATION_FINAL, ILLUMINATION_DEMODULATED, ILLUMINATION_DEMODULATED_FLOAT, ILLUMINATION_FINAL_DEMODULATED = range(4)
(ACC_PREV_ILLU, ACC_PREV_ILLU_SQUARED, ACC_PREV_DEPTH, ACC_PREV_NORMAL, ACC_SPP, ACC_PREV_SPP, ACC_MOTION, ACC_NEXT_DEPTH) = range(8)
BMFR_IMAGE_DENOISED, BMFR_IMAGE_FEATURES, BMFR_IMAGE_WEIGHTS = range(3)
PEER_HANDLE_BYTES = 64
HALO_MAX_PEERS = 8


class ImageInfo(C.Structure):
    _fields_ = [("data", C.c_void_p), ("format", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("layers", C.c_uint32), ("row_pitch", C.c_uint64), ("layer_pitch", C.c_uint64),
                ("size_bytes", C.c_uint64), ("owned", C.c_int32)]


class HaloCopy(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("src_pitch", C.c_uint64), ("dst_pitch", C.c_uint64),
                ("row_bytes", C.c_uint32), ("rows", C.c_uint32)]


class HaloExchangeDesc(C.Structure):
    _fields_ = [("copies", C.POINTER(HaloCopy)), ("n_copies", C.c_uint32),
                ("announce_flags", C.POINTER(C.c_void_p)), ("n_announce", C.c_uint32),
                ("ready_flags", C.POINTER(C.c_void_p)), ("n_ready", C.c_uint32),
                ("done_flags", C.POINTER(C.c_void_p)), ("n_done", C.c_uint32),
                ("wait_flags", C.POINTER(C.c_void_p)), ("n_wait", C.c_uint32)]


class PushConstants(C.Structure):
    """RayTracingPushConstants (source/renderModules/PipelineStructs.hpp:6-13)."""
    _fields_ = [("view_inverse", C.c_float * 16), ("proj_inverse", C.c_float * 16), ("prev_view", C.c_float * 16),
                ("frame_number", C.c_uint32), ("sample_number", C.c_uint32)]


class CameraMatrices(C.Structure):
    """CameraMatrices (source/io/RenderIO.hpp:28-34)."""
    _fields_ = [("view", C.c_float * 16), ("inv_view", C.c_float * 16), ("has_proj", C.c_int32),
                ("proj", C.c_float * 16), ("inv_proj", C.c_float * 16)]


H = C.c_void_p          # every opaque handle
PH = C.POINTER(C.c_void_p)
u32, u64, i32 = C.c_uint32, C.c_uint64, C.c_int

# name -> argtypes (restype is int unless listed in _RESTYPES)
_PROTOS = {
    "vkpbrt_context_create": [i32, C.c_void_p, PH],
    "vkpbrt_context_destroy": [H],
    "vkpbrt_context_synchronize": [H],
    "vkpbrt_context_stream": [H, PH],
    "vkpbrt_context_launch_count": [H, C.POINTER(u64)],
    "vkpbrt_debug_tonemap_sweep": [H, C.POINTER(u64), C.POINTER(u32)],
    "vkpbrt_image_create": [H, u32, u32, u32, u32, PH],
    "vkpbrt_image_wrap": [H, u32, u32, u32, u32, C.c_void_p, PH],
    "vkpbrt_image_set_data": [H, C.c_void_p],
    "vkpbrt_image_compile": [H],
    "vkpbrt_image_info_get": [H, C.POINTER(ImageInfo)],
    "vkpbrt_image_upload": [H, C.c_void_p, u64],
    "vkpbrt_image_download": [H, C.c_void_p, u64],
    "vkpbrt_image_clear": [H],
    "vkpbrt_image_copy_record": [H, H],
    "vkpbrt_mat4_inverse": [C.c_void_p, C.c_void_p],
    "vkpbrt_device_count": [C.POINTER(i32)],
    "vkpbrt_device_uuid": [i32, C.c_void_p],
    "vkpbrt_image_retain": [H],
    "vkpbrt_image_release": [H],
    "vkpbrt_gbuffer_create": [H, u32, u32, PH],
    "vkpbrt_gbuffer_create_from_images": [H, H, H, H, H, PH],
    "vkpbrt_gbuffer_compile": [H],
    "vkpbrt_gbuffer_image": [H, u32, PH],
    "vkpbrt_gbuffer_destroy": [H],
    "vkpbrt_illumination_buffer_create": [H, u32, u32, u32, PH],
    "vkpbrt_illumination_buffer_create_from_images": [H, u32, PH, u32, PH],
    "vkpbrt_illumination_buffer_compile": [H],
    "vkpbrt_illumination_buffer_type": [H, C.POINTER(u32), C.POINTER(u32)],
    "vkpbrt_illumination_buffer_image": [H, u32, PH],
    "vkpbrt_illumination_buffer_destroy": [H],
    "vkpbrt_accumulation_buffer_create": [H, u32, u32, PH],
    "vkpbrt_accumulation_buffer_compile": [H],
    "vkpbrt_accumulation_buffer_image": [H, u32, PH],
    "vkpbrt_accumulation_buffer_copy_to_back_images": [H, H, H],
    "vkpbrt_accumulation_buffer_destroy": [H],
    "vkpbrt_accumulator_create": [H, H, H, i32, i32, i32, PH],
    "vkpbrt_accumulator_compile_images": [H],
    "vkpbrt_accumulator_accumulated_illumination": [H, PH],
    "vkpbrt_accumulator_accumulation_buffer": [H, PH],
    "vkpbrt_accumulator_set_camera_matrices": [H, i32, C.POINTER(CameraMatrices), C.POINTER(CameraMatrices)],
    "vkpbrt_accumulator_record": [H],
    "vkpbrt_accumulator_set_row_range": [H, i32, i32],
    "vkpbrt_accumulator_set_force_scalar": [H, i32],
    "vkpbrt_accumulator_set_max_displacement_rows": [H, i32],
    "vkpbrt_accumulator_displacement_violations": [H, C.POINTER(u32)],
    "vkpbrt_accumulator_destroy": [H],
    "vkpbrt_bmfr_create": [H, u32, u32, u32, u32, H, H, H, u32, PH],
    "vkpbrt_bmfr_set_debug_outputs": [H, i32],
    "vkpbrt_bmfr_compile": [H],
    "vkpbrt_bmfr_set_lane": [H, i32],
    "vkpbrt_bmfr_set_position_type": [H, i32],
    "vkpbrt_bmfr_record": [H, C.POINTER(PushConstants)],
    "vkpbrt_bmfr_set_block_row_range": [H, i32, i32],
    "vkpbrt_bmfr_final_image": [H, PH],
    "vkpbrt_bmfr_image_get": [H, u32, PH],
    "vkpbrt_bmfr_destroy": [H],
    "vkpbrt_bfr_create": [H, u32, u32, u32, u32, H, H, H, PH],
    "vkpbrt_bfr_compile": [H],
    "vkpbrt_bfr_set_lane": [H, i32],
    "vkpbrt_bfr_record": [H, C.POINTER(PushConstants)],
    "vkpbrt_bfr_final_image": [H, PH],
    "vkpbrt_bfr_denoised_image": [H, PH],
    "vkpbrt_bfr_destroy": [H],
    "vkpbrt_bfr_blender_create": [H, u32, u32, H, H, H, H, H, u32, u32, u32, PH],
    "vkpbrt_bfr_blender_compile": [H],
    "vkpbrt_bfr_blender_record": [H],
    "vkpbrt_bfr_blender_final_image": [H, PH],
    "vkpbrt_bfr_blender_destroy": [H],
    "vkpbrt_taa_create": [H, u32, u32, u32, u32, H, H, H, PH],
    "vkpbrt_taa_set_fix_swizzle": [H, i32],
    "vkpbrt_taa_compile": [H],
    "vkpbrt_taa_set_force_scalar": [H, i32],
    "vkpbrt_taa_set_strip_rows": [H, i32],
    "vkpbrt_taa_record_part": [H, C.c_void_p, i32, i32, i32],
    "vkpbrt_taa_record_parts": [H, C.c_void_p, i32, i32, i32, i32, i32],
    "vkpbrt_format_converter_create": [H, H, u32, u32, u32, C.POINTER(H)],
    "vkpbrt_format_converter_compile_images": [H],
    "vkpbrt_format_converter_record": [H],
    "vkpbrt_format_converter_final_image": [H, C.POINTER(H)],
    "vkpbrt_format_converter_destroy": [H],
    "vkpbrt_demodulate_record": [H, H, H, H, H],
    "vkpbrt_gbuffer_import_record": [H, H, C.c_void_p, H, H],
    "vkpbrt_stream_create": [H, i32, C.POINTER(C.c_void_p)],
    "vkpbrt_stream_destroy": [H, C.c_void_p],
    "vkpbrt_banded_rank_create": [H, u32, u32, i32, i32, i32, i32, i32, C.c_void_p, C.c_void_p, C.c_void_p, u32, C.POINTER(H)],
    "vkpbrt_banded_rank_input_rows": [H, C.POINTER(i32), C.POINTER(i32)],
    "vkpbrt_banded_rank_block_rows": [H, C.POINTER(i32)],
    "vkpbrt_banded_rank_owned_rows": [H, u32, C.POINTER(i32), C.POINTER(i32)],
    "vkpbrt_banded_rank_bind_inputs": [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    "vkpbrt_banded_rank_run_frame": [H, u32, C.c_void_p],
    "vkpbrt_banded_rank_flush": [H],
    "vkpbrt_banded_rank_check": [H],
    "vkpbrt_banded_rank_image": [H, u32, C.POINTER(H)],
    "vkpbrt_banded_rank_stats": [H, C.POINTER(u64), C.POINTER(u64)],
    "vkpbrt_banded_rank_destroy": [H],
    "vkpbrt_taa_record": [H, C.POINTER(PushConstants)],
    "vkpbrt_taa_set_row_range": [H, i32, i32],
    "vkpbrt_taa_final_image": [H, PH],
    "vkpbrt_taa_history_image": [H, PH],
    "vkpbrt_taa_destroy": [H],
    "vkpbrt_import_external_memory_fd": [H, i32, u64, u64, u64, PH, PH],
    "vkpbrt_import_external_memory_fd_ex": [H, i32, u64, u64, u64, i32, PH, PH],
    "vkpbrt_external_memory_destroy": [H],
    "vkpbrt_import_external_semaphore_fd": [H, i32, i32, PH],
    "vkpbrt_external_semaphore_wait": [H, u64],
    "vkpbrt_external_semaphore_signal": [H, u64],
    "vkpbrt_external_semaphore_destroy": [H],
    "vkpbrt_peer_export": [H, C.c_void_p, C.c_void_p, C.POINTER(u64)],
    "vkpbrt_peer_open": [H, C.c_void_p, PH],
    "vkpbrt_peer_close": [H, C.c_void_p],
    "vkpbrt_halo_exchange_create": [H, C.c_void_p, u32, PH],
    "vkpbrt_halo_exchange_start": [H, C.c_void_p, C.c_void_p, u32],
    "vkpbrt_halo_exchange_start_gated": [H, C.c_void_p, C.c_void_p, u32, u32],
    "vkpbrt_halo_exchange_wait": [H, C.c_void_p, u32],
    "vkpbrt_halo_exchange_stats": [H, C.POINTER(u64), C.POINTER(u64), C.POINTER(u32)],
    "vkpbrt_halo_exchange_destroy": [H],
}
_RESTYPES = {"vkpbrt_last_error": C.c_char_p, "vkpbrt_version": C.c_char_p, "vkpbrt_format_texel_size": u32}
EXPORTS = sorted(list(_PROTOS) + list(_RESTYPES))

_lib = None


def lib() -> C.CDLL:
    """Loads libvkpbrt_b200.so (built in-tree by vulkanpbrt_b200.build).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise ImportError(f"{_LIB_PATH} is missing: run `python -m vulkanpbrt_b200.build` "
                              "(the package has no fallback path without its CUDA library)")
        _lib = configure(C.CDLL(str(_LIB_PATH)))
    return _lib


def configure(l: C.CDLL) -> C.CDLL:
    """declares the prototypes of include/vkpbrt_b200.h on a loaded library"""
    for name, args in _PROTOS.items():
        fn = getattr(l, name)
        fn.argtypes = args
        fn.restype = C.c_int
    l.vkpbrt_last_error.restype = C.c_char_p
    l.vkpbrt_last_error.argtypes = []
    l.vkpbrt_version.restype = C.c_char_p
    l.vkpbrt_version.argtypes = []
    l.vkpbrt_format_texel_size.restype = u32
    l.vkpbrt_format_texel_size.argtypes = [u32]
    return l


def check(rc: int) -> None:
    if rc != OK:
        msg = lib().vkpbrt_last_error()
        raise VkpbrtError(rc, msg.decode() if msg else "")


def call(name: str, *args) -> None:
    check(getattr(lib(), name)(*args))


_ZERO16 = (C.c_float * 16)()


def mat16(values) -> "C.Array":
    """16 floats as a C array.  The per-frame host path converts ~10 matrices: a float32 ndarray (what the cameras hold) is
    copied in one call instead of element by element (7 us -> 0.6 us each)."""
    if isinstance(values, (C.c_float * 16)):
        return values
    try:
        import numpy as np
        if isinstance(values, np.ndarray) and values.dtype == np.float32 and values.size == 16 and values.flags["C_CONTIGUOUS"]:
            return (C.c_float * 16).from_buffer_copy(values)
    except ImportError:     # pragma: no cover
        pass
    flat = [float(v) for v in values]
    assert len(flat) == 16
    return (C.c_float * 16)(*flat)
