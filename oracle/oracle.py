"""ctypes binding of oracle/liboracle.so -- the CPU restatement of the reference shaders.

TEST INFRASTRUCTURE ONLY (see the header of vkpbrt_oracle.c): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by the
vulkanpbrt_b200 package.  OracleChain restates the reference's HOST logic for the path: the
push-constant filling of Accumulator::set_camera_matrices (source/renderModules/Accumulator.cpp:
85-117), the dispatch order of the frame (source/VulkanPBRT.cpp:551-618) and the
end-of-frame copies (source/buffers/AccumulationBuffer.cpp:72-244, source/renderModules/Taa.cpp:106).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "liboracle.so"


class AccPush(C.Structure):
    """push-constant block of accumulator.comp:20-27"""
    _fields_ = [("view", C.c_float * 16), ("inv_view", C.c_float * 16), ("prev_view", C.c_float * 16),
                ("prev_origin", C.c_float * 4), ("frame_number", C.c_uint32)]


def build(force: bool = False) -> Path:
    src = _DIR / "vkpbrt_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        r = subprocess.run(["make", "-C", str(_DIR), "-B", "liboracle.so"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building the oracle failed:\n" + r.stdout + r.stderr)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()   # no-op when liboracle.so is newer than its source
        l = C.CDLL(str(_LIB_PATH))
        vp, i32, u32 = C.c_void_p, C.c_int, C.c_uint32
        l.vkpbrt_oracle_f32_to_f16.argtypes = [C.c_float]; l.vkpbrt_oracle_f32_to_f16.restype = C.c_uint16
        l.vkpbrt_oracle_f16_to_f32.argtypes = [C.c_uint16]; l.vkpbrt_oracle_f16_to_f32.restype = C.c_float
        l.vkpbrt_oracle_f32_to_unorm8.argtypes = [C.c_float]; l.vkpbrt_oracle_f32_to_unorm8.restype = C.c_uint8
        l.vkpbrt_oracle_mat_inverse.argtypes = [vp, vp]; l.vkpbrt_oracle_mat_inverse.restype = None
        l.vkpbrt_oracle_vsg_inverse.argtypes = [vp, vp]; l.vkpbrt_oracle_vsg_inverse.restype = None
        l.vkpbrt_oracle_mat_mul.argtypes = [vp, vp, vp]; l.vkpbrt_oracle_mat_mul.restype = None
        l.vkpbrt_oracle_accumulator.argtypes = [i32, i32, i32, C.POINTER(AccPush), vp, i32, vp, vp, vp, vp, vp, vp, vp]
        l.vkpbrt_oracle_accumulator.restype = None
        l.vkpbrt_oracle_bmfr_block_offset.argtypes = [i32, i32, u32, C.POINTER(i32), C.POINTER(i32)]
        l.vkpbrt_oracle_bmfr_block_offset.restype = None
        l.vkpbrt_oracle_bfr_block_offset.argtypes = [u32, C.POINTER(i32), C.POINTER(i32)]
        l.vkpbrt_oracle_bfr_block_offset.restype = None
        l.vkpbrt_oracle_bmfr_random.argtypes = [u32]; l.vkpbrt_oracle_bmfr_random.restype = C.c_float
        l.vkpbrt_oracle_bmfr_pre.argtypes = [i32, i32, i32, u32, vp, vp, vp, vp]; l.vkpbrt_oracle_bmfr_pre.restype = None
        l.vkpbrt_oracle_bmfr_fit.argtypes = [i32, i32, i32, i32, u32, vp, vp]; l.vkpbrt_oracle_bmfr_fit.restype = None
        l.vkpbrt_oracle_bmfr_post.argtypes = [i32, i32, i32, u32] + [vp] * 9; l.vkpbrt_oracle_bmfr_post.restype = None
        l.vkpbrt_oracle_bmfr_pre_ex.argtypes = [i32, i32, i32, u32, i32, vp, vp, vp, vp, vp, vp]; l.vkpbrt_oracle_bmfr_pre_ex.restype = None
        l.vkpbrt_oracle_bmfr_post_ex.argtypes = [i32, i32, i32, u32, i32, vp, vp] + [vp] * 9; l.vkpbrt_oracle_bmfr_post_ex.restype = None
        l.vkpbrt_oracle_bfr.argtypes = [i32, i32, i32, u32] + [vp] * 8; l.vkpbrt_oracle_bfr.restype = None
        l.vkpbrt_oracle_bfr_lr.argtypes = [i32]; l.vkpbrt_oracle_bfr_lr.restype = C.c_float
        l.vkpbrt_oracle_bfr_blender.argtypes = [i32, i32, i32] + [vp] * 6; l.vkpbrt_oracle_bfr_blender.restype = None
        l.vkpbrt_oracle_taa.argtypes = [i32, i32, u32, i32] + [vp] * 4; l.vkpbrt_oracle_taa.restype = None
        l.vkpbrt_oracle_format_converter.argtypes = [i32, i32, i32, vp, vp]; l.vkpbrt_oracle_format_converter.restype = None
        l.vkpbrt_oracle_gbuffer_import.argtypes = [i32, i32] + [vp] * 7; l.vkpbrt_oracle_gbuffer_import.restype = None
        l.vkpbrt_oracle_gbuffer_export.argtypes = [i32, i32] + [vp] * 8; l.vkpbrt_oracle_gbuffer_export.restype = None
        l.vkpbrt_oracle_demodulate.argtypes = [i32, i32, vp, vp, vp, vp]; l.vkpbrt_oracle_demodulate.restype = None
        l.vkpbrt_oracle_num_threads.argtypes = []; l.vkpbrt_oracle_num_threads.restype = i32
        l.vkpbrt_oracle_set_num_threads.argtypes = [i32]; l.vkpbrt_oracle_set_num_threads.restype = None
        _lib = l
    return _lib


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"], "oracle arrays must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)


def f32_to_f16_bits(a: np.ndarray) -> np.ndarray:
    """the oracle's fp32->fp16 (RTE); numpy's astype(float16) is the same rounding"""
    return np.asarray(a, dtype=np.float32).astype(np.float16).view(np.uint16)


def f16_bits_to_f32(a: np.ndarray) -> np.ndarray:
    return np.asarray(a, dtype=np.uint16).view(np.float16).astype(np.float32)


def mat_inverse(m) -> np.ndarray:
    a = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    out = np.empty(16, np.float32)
    lib().vkpbrt_oracle_mat_inverse(_p(a), _p(out))
    return out


def vsg_inverse(m) -> np.ndarray:
    """vsg::inverse(mat4) as the reference's HOST code calls it (not the shader's inverse(): that is mat_inverse)"""
    a = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    out = np.empty(16, np.float32)
    lib().vkpbrt_oracle_vsg_inverse(_p(a), _p(out))
    return out


def bmfr_block_offset(b: int, frame: int):
    ox, oy = C.c_int(), C.c_int()
    lib().vkpbrt_oracle_bmfr_block_offset(b, b, frame, C.byref(ox), C.byref(oy))
    return ox.value, oy.value


def bfr_block_offset(frame: int):
    ox, oy = C.c_int(), C.c_int()
    lib().vkpbrt_oracle_bfr_block_offset(frame, C.byref(ox), C.byref(oy))
    return ox.value, oy.value


IDENTITY = np.eye(4, dtype=np.float32).reshape(16)


class OracleChain:
    """Reference frame on the CPU: accumulator -> (BMFR | BFR x1 | BFR x3 + blender) -> [TAA] -> copy back."""

    def __init__(self, width: int, height: int, denoiser: str = "bmfr", block: int = 32, use_taa: bool = False,
                 separate_matrices: bool = True, raw_f16: bool = False, fix_taa_swizzle: bool = False,
                 blend_radius: int = 2, position_type: int = 0):
        W, H = width, height
        self.W, self.H, self.denoiser, self.block = W, H, denoiser, block
        self.use_taa, self.separate, self.raw_f16, self.fix_swz = use_taa, separate_matrices, raw_f16, fix_taa_swizzle
        self.blend_radius = blend_radius
        self.position_type = position_type      # bmfrGeneral.comp:30-31 POSITION_TYPE (BMFR only)
        z = np.zeros
        # AccumulationBuffer (AccumulationBuffer.cpp:245-339) + accumulated illumination; zero-initialised
        self.prev_depth = z((H, W), np.float32)
        self.prev_illu = z((H, W, 4), np.uint16)
        self.prev_spp = z((H, W), np.uint8)
        self.spp = z((H, W), np.uint8)
        self.motion = z((H, W, 2), np.uint16)
        self.illum = z((H, W, 4), np.uint16)
        self.average_squared = z((H, W, 4), np.uint16)
        blocks = [8, 16, 32] if denoiser.endswith("x3") else [block]
        self.blocks = blocks
        self.denoised = {b: z((2, H, W, 4), np.uint16) for b in blocks}
        self.finals = {b: z((H, W, 4), np.uint8) for b in blocks}
        self.features = None
        self.weights = None
        self.blend_final = z((H, W, 4), np.uint8)
        self.taa_final = z((H, W, 4), np.uint8)
        self.taa_history = z((H, W, 4), np.uint8)
        self.prev_view = IDENTITY.copy()          # RayTracingPushConstants.prev_view, default-constructed
        self.prev_cam = None
        self.pc = AccPush()

    # Accumulator::set_camera_matrices, Accumulator.cpp:85-117
    def _set_camera_matrices(self, frame_index, cam):
        pc = self.pc
        put = lambda dst, src: [dst.__setitem__(i, float(v)) for i, v in enumerate(src)]
        if self.separate:
            put(pc.view, cam.inv_proj)
            put(pc.inv_view, cam.inv_view)
            if frame_index != 0:
                put(pc.prev_view, self.prev_view)
                inv = vsg_inverse(self.prev_view)               # :100 inverse(prev.view)[3]: vsg's host-side inverse
                put(pc.prev_origin, [inv[12], inv[13], inv[14], 1.0])
        else:
            from vulkanpbrt_b200.pipeline import _combined   # host-side matrix prep shared with the pipeline helper
            vp, ivp = _combined(cam)
            put(pc.view, vp)
            put(pc.inv_view, ivp)
            if frame_index != 0:
                pvp, pivp = _combined(self.prev_cam)
                put(pc.prev_view, pvp)
                inv_w = np.float32(1.0) / np.float32(pivp[11])     # :110-111 vsg's vec4 /= multiplies by the reciprocal
                put(pc.prev_origin, [np.float32(pivp[8 + i]) * inv_w for i in range(4)])
        pc.frame_number = frame_index

    def final(self) -> np.ndarray:
        if self.use_taa:
            return self.taa_final
        if self.denoiser.endswith("x3"):
            return self.blend_final
        return self.finals[self.block]

    def denoiser_final(self) -> np.ndarray:
        return self.blend_final if self.denoiser.endswith("x3") else self.finals[self.block]

    def run_frame(self, frame_index: int, frame, keep_debug: bool = False, motion_override=None) -> None:
        """motion_override: a motion plane (rg16f bits) that replaces the accumulator's before the denoisers run -- what a
        caller that uploads its own motion vectors does (the image is public: AccumulationBuffer::motion)"""
        L, W, H = lib(), self.W, self.H
        self._set_camera_matrices(frame_index, frame.camera)
        if self.raw_f16:
            src = np.ascontiguousarray(frame.illumination.astype(np.float16).view(np.uint16))
        else:
            src = np.ascontiguousarray(frame.illumination, dtype=np.float32)
        depth = np.ascontiguousarray(frame.depth, dtype=np.float32)
        normal = np.ascontiguousarray(frame.normal, dtype=np.float32)
        albedo = np.ascontiguousarray(frame.albedo, dtype=np.uint8)
        L.vkpbrt_oracle_accumulator(W, H, 1 if self.separate else 0, C.byref(self.pc), _p(src), 1 if self.raw_f16 else 0,
                                    _p(depth), _p(self.prev_depth), _p(self.prev_illu), _p(self.prev_spp),
                                    _p(self.motion), _p(self.spp), _p(self.illum))
        if motion_override is not None:
            self.motion[...] = motion_override
        for b in self.blocks:
            if self.denoiser.startswith("bmfr"):
                T = 64 if b == 8 else 256
                Wb, Hb = W // b + 2, H // b + 2
                feat = np.zeros((13, Hb * b, Wb * b), np.uint16)
                wts = np.zeros((30, Hb, Wb), np.float32)
                iv = np.ascontiguousarray(frame.camera.inv_view, dtype=np.float32)      # RayTracingPushConstants (VulkanPBRT.cpp:561)
                ip = np.ascontiguousarray(frame.camera.inv_proj, dtype=np.float32)
                L.vkpbrt_oracle_bmfr_pre_ex(W, H, b, frame_index, self.position_type, _p(iv), _p(ip), _p(self.illum), _p(depth), _p(normal),
                                            _p(feat))
                L.vkpbrt_oracle_bmfr_fit(W, H, b, T, frame_index, _p(feat), _p(wts))
                L.vkpbrt_oracle_bmfr_post_ex(W, H, b, frame_index, self.position_type, _p(iv), _p(ip), _p(self.illum), _p(depth), _p(normal),
                                             _p(albedo), _p(self.motion), _p(self.spp), _p(wts), _p(self.denoised[b]), _p(self.finals[b]))
                if keep_debug:
                    self.features, self.weights = feat, wts
            else:
                L.vkpbrt_oracle_bfr(W, H, b, frame_index, _p(self.illum), _p(depth), _p(normal), _p(albedo),
                                    _p(self.motion), _p(self.spp), _p(self.denoised[b]), _p(self.finals[b]))
        if self.denoiser.endswith("x3"):
            L.vkpbrt_oracle_bfr_blender(W, H, self.blend_radius, _p(self.illum), _p(self.average_squared),
                                        _p(self.finals[8]), _p(self.finals[16]), _p(self.finals[32]), _p(self.blend_final))
        if self.use_taa:
            L.vkpbrt_oracle_taa(W, H, frame_index, 1 if self.fix_swz else 0, _p(self.motion), _p(self.denoiser_final()),
                                _p(self.taa_history), _p(self.taa_final))
            self.taa_history[...] = self.taa_final          # Taa.cpp:106 raw copy
        # AccumulationBuffer::copy_to_back_images (AccumulationBuffer.cpp:72-244)
        self.prev_depth[...] = depth
        self.prev_spp[...] = self.spp
        self.prev_illu[...] = self.illum
        self.prev_view = np.asarray(frame.camera.view, dtype=np.float32).copy()     # VulkanPBRT.cpp:591
        self.prev_cam = frame.camera


def format_converter(src: np.ndarray) -> np.ndarray:
    """formatConverter.comp on a [H][W][4] float32 / float16-bits (uint16) / uint8 image -> [H][W][4] BGRA8"""
    H, W = src.shape[:2]
    fmt = {np.dtype(np.float32): 0, np.dtype(np.uint16): 1, np.dtype(np.uint8): 2}[src.dtype]
    out = np.zeros((H, W, 4), np.uint8)
    lib().vkpbrt_oracle_format_converter(W, H, fmt, _p(np.ascontiguousarray(src)), _p(out))
    return out


def demodulate(radiance: np.ndarray, albedo: np.ndarray, position_x: np.ndarray) -> np.ndarray:
    H, W = radiance.shape[:2]
    out = np.zeros((H, W, 4), np.float32)
    lib().vkpbrt_oracle_demodulate(W, H, _p(np.ascontiguousarray(radiance, np.float32)), _p(np.ascontiguousarray(albedo, np.float32)),
                                   _p(np.ascontiguousarray(position_x, np.float32)), _p(out))
    return out


def gbuffer_import(inv_view, position=None, normal=None, albedo=None):
    """GBufferIO::import_g_buffer_position's conversions; returns (depth, normal (theta, phi), albedo rgba8), None where no input"""
    ref = next(a for a in (position, normal, albedo) if a is not None)
    H, W = ref.shape[:2]
    c = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    position, normal, albedo = c(position), c(normal), c(albedo)
    d = np.zeros((H, W), np.float32) if position is not None else None
    n = np.zeros((H, W, 2), np.float32) if normal is not None else None
    a = np.zeros((H, W, 4), np.uint8) if albedo is not None else None
    iv = np.ascontiguousarray(inv_view, np.float32) if inv_view is not None else None
    q = lambda x: None if x is None else _p(x)
    lib().vkpbrt_oracle_gbuffer_import(W, H, q(iv), q(position), q(normal), q(albedo), q(d), q(n), q(a))
    return d, n, a


def gbuffer_export(inv_view=None, inv_proj=None, depth=None, normal=None, unorm=None):
    """GBufferIO::export_g_buffer's conversions (RenderIO.cpp:312-382); returns (position, cartesian normal, float rgba), each
    rgba32f, None where there is no input (or, for positions, no separate projection matrix)"""
    ref = next(a for a in (depth, normal, unorm) if a is not None)
    H, W = ref.shape[:2]
    f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
    depth, normal, inv_view, inv_proj = f(depth), f(normal), f(inv_view), f(inv_proj)
    unorm = None if unorm is None else np.ascontiguousarray(unorm, np.uint8)
    p = np.zeros((H, W, 4), np.float32) if depth is not None and inv_proj is not None and inv_view is not None else None
    n = np.zeros((H, W, 4), np.float32) if normal is not None else None
    u = np.zeros((H, W, 4), np.float32) if unorm is not None else None
    q = lambda x: None if x is None else _p(x)
    lib().vkpbrt_oracle_gbuffer_export(W, H, q(inv_view), q(inv_proj), q(depth), q(normal), q(unorm), q(p), q(n), q(u))
    return p, n, u
