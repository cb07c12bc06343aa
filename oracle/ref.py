"""oracle/_ref driver: runs the reference's OWN shader source (compiled through oracle/glsl_shim) on the CPU.

TEST INFRASTRUCTURE.  RefChain replays one frame exactly like the reference's command list: the descriptor bindings
of denoisers/BMFR.hpp:28-30, BFR.hpp:24-36, BFRBlender.hpp:20-21, Taa.hpp:11, accumulator.comp:3-18, the dispatch
sizes of Accumulator.cpp:80-81, BMFR.cpp:208-209, BFR.cpp:134, BFRBlender.cpp:88-89, Taa.cpp:104-105 and the push
constant blocks.  It shares the plane containers with OracleChain so the two can be compared field by field.
"""
from __future__ import annotations

import ctypes as C
import math
import subprocess
from pathlib import Path

import numpy as np

from . import oracle as O

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / "_ref" / "libref.so"
F_R32F, F_RG32F, F_RGBA8, F_BGRA8, F_RG16F, F_R8, F_RGBA16F, F_RGBA32F, F_R16F = range(1, 10)


class _Binding(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("layers", C.c_int), ("format", C.c_int)]


class _RtPush(C.Structure):    # RayTracingPushConstants, source/renderModules/PipelineStructs.hpp:6-13
    _fields_ = [("view_inverse", C.c_float * 16), ("proj_inverse", C.c_float * 16), ("prev_view", C.c_float * 16),
                ("frame_number", C.c_uint32), ("sample_number", C.c_uint32)]


def build(reference: str = "/root/reference") -> bool:
    """compiles oracle/_ref from the reference tree when it is mounted; returns availability"""
    shaders = Path(reference, "shaders")
    if shaders.exists():
        # the Makefile deletes its intermediates (text derived from the reference is never kept), so make itself would
        # rebuild everything on every call: decide here whether the library is older than anything it is made from
        deps = [p for p in (_DIR / "glsl_shim").iterdir() if p.is_file()] + [p for p in shaders.iterdir() if p.is_file()]
        if not _LIB.exists() or any(p.stat().st_mtime > _LIB.stat().st_mtime for p in deps):
            r = subprocess.run(["make", "-C", str(_DIR / "glsl_shim"), f"REF={reference}"], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("oracle/_ref build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return _LIB.exists()


def available() -> bool:
    return _LIB.exists()


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(str(_LIB))
        l.ref_dispatch.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(_Binding), C.c_int]
        l.ref_dispatch.restype = C.c_int
        l.ref_available.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int]
        _lib = l
    return _lib


def _bind(entries):
    """entries: {binding: (array, format)}; array shape [layers?][H][W][..]"""
    n = max(entries) + 1
    arr = (_Binding * n)()
    for b, (a, fmt) in entries.items():
        assert a.flags["C_CONTIGUOUS"]
        chan = {F_R32F: 1, F_RG32F: 2, F_RGBA8: 4, F_BGRA8: 4, F_RG16F: 2, F_R8: 1, F_RGBA16F: 4, F_RGBA32F: 4, F_R16F: 1}[fmt]
        shape = a.shape[:-1] if chan > 1 else a.shape
        layers, h, w = (shape if len(shape) == 3 else (1,) + tuple(shape))
        arr[b] = _Binding(a.ctypes.data, w, h, layers, fmt)
    return arr, n


def dispatch(shader: str, key, W, H, radius, push, gx, gy, entries) -> None:
    arr, n = _bind(entries)
    rc = lib().ref_dispatch(shader.encode(), key[0], key[1], key[2], W, H, radius, C.byref(push) if push is not None else None,
                            C.sizeof(push) if push is not None else 0, gx, gy, arr, n)
    if rc != 0:
        raise RuntimeError(f"oracle/_ref: shader {shader}{key} not built (rc={rc})")


class RefChain(O.OracleChain):
    """the reference frame executed by the reference's shader source"""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        H, W = self.H, self.W
        self.material = np.zeros((H, W, 4), np.uint8)
        self.prev_normal = np.zeros((H, W, 2), np.float32)
        self.prev_illu_squared = np.zeros((H, W, 4), np.uint16)
        self.illum_squared = np.zeros((H, W, 4), np.uint16)
        self.rt = _RtPush()

    def run_frame(self, frame_index: int, frame, keep_debug: bool = False) -> None:
        W, H = self.W, self.H
        self._set_camera_matrices(frame_index, frame.camera)
        src = (np.ascontiguousarray(frame.illumination.astype(np.float16).view(np.uint16)) if self.raw_f16
               else np.ascontiguousarray(frame.illumination, dtype=np.float32))
        depth = np.ascontiguousarray(frame.depth, dtype=np.float32)
        normal = np.ascontiguousarray(frame.normal, dtype=np.float32)
        albedo = np.ascontiguousarray(frame.albedo, dtype=np.uint8)
        # ---- accumulator.comp (Accumulator.cpp:72-83) ----
        dispatch("accumulator_sep" if self.separate else "accumulator", (16, 16, 0), W, H, 0, self.pc, math.ceil(W / 16), math.ceil(H / 16), {
            0: (src, F_RGBA16F if self.raw_f16 else F_RGBA32F), 1: (depth, F_R32F), 2: (normal, F_RG32F), 3: (self.material, F_RGBA8),
            4: (albedo, F_RGBA8), 5: (self.prev_depth, F_R32F), 6: (self.prev_normal, F_RG32F), 7: (self.motion, F_RG16F),
            8: (self.spp, F_R8), 9: (self.prev_spp, F_R8), 10: (self.prev_illu, F_RGBA16F), 11: (self.illum, F_RGBA16F),
            12: (self.prev_illu_squared, F_RGBA16F), 13: (self.illum_squared, F_RGBA16F)})
        # RayTracingPushConstants (VulkanPBRT.cpp:561-563); only frameNumber is live for POSITION_DEPTH
        for i in range(16):
            self.rt.view_inverse[i] = float(frame.camera.inv_view[i])
            self.rt.proj_inverse[i] = float(frame.camera.inv_proj[i])
            self.rt.prev_view[i] = float(self.prev_view[i])
        self.rt.frame_number = frame_index
        self.rt.sample_number = 0
        for b in self.blocks:
            common = {0: (depth, F_R32F), 1: (normal, F_RG32F), 2: (self.material, F_RGBA8), 3: (albedo, F_RGBA8), 4: (self.motion, F_RG16F),
                      5: (self.spp, F_R8), 6: (self.denoised[b], F_RGBA16F), 7: (self.finals[b], F_BGRA8), 8: (self.illum, F_RGBA16F),
                      9: (self.denoised[b], F_RGBA16F)}
            Wb, Hb = W // b + 2, H // b + 2
            if self.denoiser.startswith("bmfr"):
                T = 64 if b == 8 else 256
                feat = np.zeros((13, Hb * b, Wb * b), np.uint16)       # BMFR.cpp:96-113
                wts = np.zeros((30, Hb, Wb), np.float32)               # BMFR.cpp:116-133
                ent = dict(common)
                ent[10] = (feat, F_R16F)
                ent[11] = (wts, F_R32F)
                sfx = {0: "", 1: "_w1", 2: "_w2"}[self.position_type]                # POSITION_TYPE instantiation (oracle/glsl_shim/Makefile)
                dispatch("bmfrPre" + sfx, (b, b, b), W, H, 0, self.rt, Wb, Hb, ent)       # BMFR.cpp:203-230: pre, fit, post
                dispatch("bmfrFit", (T, 1, b), W, H, 0, self.rt, Wb, Hb, ent)
                dispatch("bmfrPost" + sfx, (b, b, b), W, H, 0, self.rt, Wb, Hb, ent)
                if keep_debug:
                    self.features, self.weights = feat, wts
            else:
                dispatch("bfr", (b, b, 0), W, H, 0, self.rt, Wb, Hb, common)         # BFR.cpp:128-138
        if self.denoiser.endswith("x3"):
            # BFRBlender(width, height, illumination_images[0], illumination_images[1], bfr8, bfr16, bfr32) (DenoiserUtils.cpp:53-55)
            dispatch("bfrBlender", (16, 16, 0), W, H, self.blend_radius, None, math.ceil(W / 16), math.ceil(H / 16), {
                0: (self.illum, F_RGBA16F), 1: (self.average_squared, F_RGBA16F), 2: (self.finals[8], F_BGRA8), 3: (self.finals[16], F_BGRA8),
                4: (self.finals[32], F_BGRA8), 5: (self.blend_final, F_BGRA8)})
        if self.use_taa:
            if self.fix_swz:
                raise NotImplementedError("the reference has no swizzle fix")
            dispatch("taa", (16, 16, 0), W, H, 0, self.rt, math.ceil(W / 16), math.ceil(H / 16), {
                0: (self.motion, F_RG16F), 1: (self.denoiser_final(), F_BGRA8), 2: (self.taa_final, F_BGRA8), 3: (self.taa_history, F_RGBA8)})
            self.taa_history[...] = self.taa_final          # Taa.cpp:106: raw vkCmdCopyImage BGRA8 -> RGBA8
        # AccumulationBuffer::copy_to_back_images (AccumulationBuffer.cpp:72-244)
        self.prev_depth[...] = depth
        self.prev_normal[...] = normal
        self.prev_spp[...] = self.spp
        self.prev_illu[...] = self.illum
        self.prev_illu_squared[...] = self.illum_squared
        self.prev_view = np.asarray(frame.camera.view, dtype=np.float32).copy()
        self.prev_cam = frame.camera


def format_converter(src: np.ndarray) -> np.ndarray:
    """the reference's formatConverter.comp (FormatConverter.cpp:20-35: local size 16 x 16, FORMAT rgba8 on a BGRA8 view)
    on a [H][W][4] float32 / float16-bits (uint16) / uint8 image"""
    H, W = src.shape[:2]
    fmt = {np.dtype(np.float32): F_RGBA32F, np.dtype(np.uint16): F_RGBA16F, np.dtype(np.uint8): F_RGBA8}[src.dtype]
    out = np.zeros((H, W, 4), np.uint8)
    dispatch("formatConverter", (16, 16, 0), W, H, 0, None, math.ceil(W / 16), math.ceil(H / 16),
             {0: (np.ascontiguousarray(src), fmt), 1: (out, F_BGRA8)})
    return out


def demodulate(radiance: np.ndarray, albedo: np.ndarray, position_x: np.ndarray) -> np.ndarray:
    """the radiance clamp and the DEMOD_ILLUMINATION_FLOAT block of the reference's ray-generation shader
    (ptRaygen.rgen:81-88), cut out of its text and wrapped in a compute main() by glsl_shim/extract_rgen.py, on
    [H][W][4] float32 radiance / albedo planes and a [H][W] plane of the primary hit's position.x (infinite for misses)"""
    H, W = radiance.shape[:2]
    out = np.zeros((H, W, 4), np.float32)
    dispatch("demodRgen", (16, 16, 0), W, H, 0, None, math.ceil(W / 16), math.ceil(H / 16),
             {0: (np.ascontiguousarray(radiance, np.float32), F_RGBA32F), 1: (np.ascontiguousarray(albedo, np.float32), F_RGBA32F),
              2: (np.ascontiguousarray(position_x, np.float32), F_R32F), 3: (out, F_RGBA32F)})
    return out


# ---- the reference's HOST conversion code (source/io/RenderIO.cpp), oracle/host_shim ---------------------------------
_HOST_LIB = _DIR / "_ref" / "libhostref.so"
_host = None


def build_host(reference: str = "/root/reference") -> bool:
    if Path(reference, "source", "io", "RenderIO.cpp").exists():
        r = subprocess.run(["make", "-C", str(_DIR / "host_shim"), f"REF={reference}"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref host build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return _HOST_LIB.exists()


def gbuffer_import(inv_view, position=None, normal=None, albedo=None):
    """GBufferIO::convert_normal_to_spherical, GBufferIO::compress_albedo and the position -> depth block of
    import_g_buffer_position (RenderIO.cpp:101-120, :160-211) as the reference's own C++ text, compiled against vsg's
    maths headers; same interface as oracle.gbuffer_import"""
    global _host
    if _host is None:
        _host = C.CDLL(str(_HOST_LIB))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    d = n = a = None
    if position is not None:
        position = np.ascontiguousarray(position, np.float32)
        H, W = position.shape[:2]
        d = np.zeros((H, W), np.float32)
        iv = np.ascontiguousarray(inv_view, np.float32)
        assert _host.hostref_depth(p(position), W, H, p(iv), p(d)) == 0
    if normal is not None:
        normal = np.ascontiguousarray(normal, np.float32)
        H, W = normal.shape[:2]
        n = np.zeros((H, W, 2), np.float32)
        assert _host.hostref_normals(p(normal), W, H, p(n)) == 0
    if albedo is not None:
        albedo = np.ascontiguousarray(albedo, np.float32)
        H, W = albedo.shape[:2]
        a = np.zeros((H, W, 4), np.uint8)
        assert _host.hostref_albedo(p(albedo), W, H, p(a)) == 0
    return d, n, a


def gbuffer_export(matrices64=None, has_proj=True, depth=None, normal=None, unorm=None):
    """GBufferIO::depth_to_position, spherical_to_cartesian and unorm_to_float (RenderIO.cpp:312-382) as the reference's own
    C++ text.  matrices64 = view, inv_view, proj, inv_proj (16 floats each, column-major).  Returns (position, cartesian
    normal, float rgba); position is None when the reference returns no data (no separate projection matrix)"""
    global _host
    if _host is None:
        _host = C.CDLL(str(_HOST_LIB))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    pos = n = u = None
    if depth is not None:
        depth = np.ascontiguousarray(depth, np.float32)
        H, W = depth.shape
        pos = np.zeros((H, W, 4), np.float32)
        m = np.ascontiguousarray(matrices64, np.float32)
        if _host.hostref_depth_to_position(p(depth), W, H, p(m), 1 if has_proj else 0, p(pos)) != 0:
            pos = None
    if normal is not None:
        normal = np.ascontiguousarray(normal, np.float32)
        H, W = normal.shape[:2]
        n = np.zeros((H, W, 4), np.float32)
        assert _host.hostref_spherical_to_cartesian(p(normal), W, H, p(n)) == 0
    if unorm is not None:
        unorm = np.ascontiguousarray(unorm, np.uint8)
        H, W = unorm.shape[:2]
        u = np.zeros((H, W, 4), np.float32)
        assert _host.hostref_unorm_to_float(p(unorm), W, H, p(u)) == 0
    return pos, n, u
