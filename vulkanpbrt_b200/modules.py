"""Host-side mirror of the reference's render-module API for the denoising path.

Same class names, constructor argument order and method names as the reference's C++ classes
(SURVEY.md section 8(b)); every method forwards to the C ABI of include/vkpbrt_b200.h, which
launches the sm_100a kernels.  vsg types map as follows:

    vsg::ref_ptr<vsg::DescriptorImage>  -> DescriptorImage   (device image handle)
    vsg::Context&                       -> Context           (device + CUDA stream)
    vsg::ref_ptr<vsg::Commands>         -> Commands          (recorded once, replayed per frame)
    vsg::ref_ptr<vsg::PushConstants>    -> PushConstants     (shared struct, mutated per frame)

Reference: source/renderModules/{Accumulator,Taa}.hpp, source/renderModules/denoisers/{BMFR,BFR,
BFRBlender}.hpp, source/buffers/*.hpp, source/util/DenoiserUtils.cpp.
"""
from __future__ import annotations

import ctypes as C
from enum import Enum
from typing import Callable, List, Optional

import numpy as np

from . import _capi as capi
from ._capi import CameraMatrices as _CCameraMatrices
from ._capi import PushConstants as _CPushConstants
from ._capi import VkpbrtError

_NP_LAYOUT = {
    capi.FORMAT_R32_SFLOAT: (np.float32, 1),
    capi.FORMAT_R32G32_SFLOAT: (np.float32, 2),
    capi.FORMAT_R8G8B8A8_UNORM: (np.uint8, 4),
    capi.FORMAT_B8G8R8A8_UNORM: (np.uint8, 4),
    capi.FORMAT_R16G16_SFLOAT: (np.uint16, 2),
    capi.FORMAT_R8_UNORM: (np.uint8, 1),
    capi.FORMAT_R16G16B16A16_SFLOAT: (np.uint16, 4),
    capi.FORMAT_R32G32B32A32_SFLOAT: (np.float32, 4),
    capi.FORMAT_R16_SFLOAT: (np.uint16, 1),
}


class Context:
    """Stands in for vsg::Context: the device and the CUDA stream every module records onto."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = C.c_void_p()
        capi.call("vkpbrt_context_create", int(device), C.c_void_p(stream) if stream else None, C.byref(self._h))
        self.device = device

    @property
    def handle(self):
        return self._h

    def synchronize(self) -> None:
        capi.call("vkpbrt_context_synchronize", self._h)

    def wait_for_completion(self) -> None:   # vsg::Context::waitForCompletion
        self.synchronize()

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        capi.call("vkpbrt_context_stream", self._h, C.byref(s))
        return s.value or 0

    @property
    def launch_count(self) -> int:
        n = C.c_uint64()
        capi.call("vkpbrt_context_launch_count", self._h, C.byref(n))
        return n.value

    def close(self) -> None:
        if self._h:
            capi.lib().vkpbrt_context_destroy(self._h)
            self._h = C.c_void_p()


class DescriptorImage:
    """A device image (pitch-linear, tightly packed).  16-bit float formats are exposed to numpy as
    their uint16 bit patterns so comparisons with the oracle are exact."""

    def __init__(self, ctx: Context, handle: C.c_void_p, owner: bool):
        self.ctx = ctx
        self._h = handle
        self._owner = owner

    @classmethod
    def create(cls, ctx: Context, fmt: int, width: int, height: int, layers: int = 1) -> "DescriptorImage":
        h = C.c_void_p()
        capi.call("vkpbrt_image_create", ctx.handle, fmt, width, height, layers, C.byref(h))
        return cls(ctx, h, True)

    @classmethod
    def wrap(cls, ctx: Context, fmt: int, width: int, height: int, device_ptr: int, layers: int = 1) -> "DescriptorImage":
        h = C.c_void_p()
        capi.call("vkpbrt_image_wrap", ctx.handle, fmt, width, height, layers, C.c_void_p(device_ptr), C.byref(h))
        return cls(ctx, h, True)

    @property
    def handle(self):
        return self._h

    def compile(self, context: Optional[Context] = None) -> None:
        capi.call("vkpbrt_image_compile", self._h)

    def set_data(self, device_ptr: int) -> None:
        capi.call("vkpbrt_image_set_data", self._h, C.c_void_p(device_ptr))

    def info(self) -> capi.ImageInfo:
        i = capi.ImageInfo()
        capi.call("vkpbrt_image_info_get", self._h, C.byref(i))
        return i

    @property
    def device_ptr(self) -> int:
        return self.info().data or 0

    def shape_dtype(self):
        i = self.info()
        dt, ch = _NP_LAYOUT[i.format]
        shape = [i.height, i.width]
        if ch > 1:
            shape.append(ch)
        if i.layers > 1:
            shape.insert(0, i.layers)
        return tuple(shape), dt

    def upload(self, array: np.ndarray, sync: bool = True) -> None:
        shape, dt = self.shape_dtype()
        a = np.ascontiguousarray(array)
        if a.dtype == np.float16 and dt == np.uint16:
            a = a.view(np.uint16)
        if a.dtype != dt or a.size != int(np.prod(shape)):
            raise ValueError(f"upload: expected {shape} {dt}, got {a.shape} {a.dtype}")
        capi.call("vkpbrt_image_upload", self._h, a.ctypes.data_as(C.c_void_p), a.nbytes)
        if sync:
            self.ctx.synchronize()   # `a` may be a temporary

    def upload_ptr(self, host_ptr: int, nbytes: int) -> None:
        """async upload from caller-managed (ideally pinned) host memory"""
        capi.call("vkpbrt_image_upload", self._h, C.c_void_p(host_ptr), nbytes)

    def download_ptr(self, host_ptr: int, nbytes: int) -> None:
        capi.call("vkpbrt_image_download", self._h, C.c_void_p(host_ptr), nbytes)

    def download(self) -> np.ndarray:
        shape, dt = self.shape_dtype()
        out = np.empty(shape, dtype=dt)
        capi.call("vkpbrt_image_download", self._h, out.ctypes.data_as(C.c_void_p), out.nbytes)
        self.ctx.synchronize()
        return out

    def clear(self) -> None:
        capi.call("vkpbrt_image_clear", self._h)

    @property
    def __cuda_array_interface__(self):
        """lets torch/cupy view the CURRENT device buffer without a copy (multi-GPU halo exchange)"""
        shape, dt = self.shape_dtype()
        return {"shape": shape, "typestr": np.dtype(dt).str, "data": (self.device_ptr, False), "version": 2}

    def byte_view(self) -> "_ByteView":
        """the CURRENT device buffer as raw bytes, shape [layers?][H][row_pitch] uint8 (any NCCL-friendly dtype)"""
        i = self.info()
        shape = (i.height, int(i.row_pitch)) if i.layers == 1 else (i.layers, i.height, int(i.row_pitch))
        return _ByteView(i.data or 0, shape)

    def __del__(self):
        try:
            if self._owner and self._h:
                capi.lib().vkpbrt_image_release(self._h)
        except Exception:
            pass


class _ByteView:
    def __init__(self, ptr: int, shape):
        self.ptr, self.shape = ptr, tuple(shape)

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "|u1", "data": (self.ptr, False), "version": 2}


def _borrow(ctx: Context, getter: str, owner_handle, *args) -> DescriptorImage:
    h = C.c_void_p()
    capi.call(getter, owner_handle, *args, C.byref(h))
    return DescriptorImage(ctx, h, False)


# ---------------------------------------------------------------------------------------------------
# buffer bundles
# ---------------------------------------------------------------------------------------------------
class GBuffer:
    """source/buffers/GBuffer.hpp:12-21 -- public members depth, normal, material, albedo."""

    def __init__(self, ctx: Context, width: int, height: int):
        self.ctx, self.width, self.height = ctx, width, height
        self._h = C.c_void_p()
        capi.call("vkpbrt_gbuffer_create", ctx.handle, width, height, C.byref(self._h))
        self.depth = _borrow(ctx, "vkpbrt_gbuffer_image", self._h, capi.GBUFFER_DEPTH)
        self.normal = _borrow(ctx, "vkpbrt_gbuffer_image", self._h, capi.GBUFFER_NORMAL)
        self.material = _borrow(ctx, "vkpbrt_gbuffer_image", self._h, capi.GBUFFER_MATERIAL)
        self.albedo = _borrow(ctx, "vkpbrt_gbuffer_image", self._h, capi.GBUFFER_ALBEDO)

    @classmethod
    def create(cls, ctx: Context, width: int, height: int) -> "GBuffer":
        return cls(ctx, width, height)

    @classmethod
    def from_images(cls, ctx: Context, depth: DescriptorImage, normal: DescriptorImage,
                    material: Optional[DescriptorImage], albedo: DescriptorImage) -> "GBuffer":
        self = cls.__new__(cls)
        i = depth.info()
        self.ctx, self.width, self.height = ctx, i.width, i.height
        self._h = C.c_void_p()
        capi.call("vkpbrt_gbuffer_create_from_images", ctx.handle, depth.handle, normal.handle,
                  material.handle if material else None, albedo.handle, C.byref(self._h))
        self.depth, self.normal, self.material, self.albedo = depth, normal, material, albedo
        return self

    @property
    def handle(self):
        return self._h

    def compile(self, context: Optional[Context] = None) -> None:
        capi.call("vkpbrt_gbuffer_compile", self._h)

    def update_image_layouts(self, context: Optional[Context] = None) -> None:
        """Vulkan image layouts do not exist for linear device memory: kept for call-site parity."""

    def __del__(self):
        try:
            capi.lib().vkpbrt_gbuffer_destroy(self._h)
        except Exception:
            pass


class IlluminationBuffer:
    """source/buffers/IlluminationBuffer.hpp:14-29 -- illumination_images[i]."""
    TYPE = capi.ILLUMINATION_FINAL

    def __init__(self, ctx: Context, width: int, height: int, _handle=None):
        self.ctx, self.width, self.height = ctx, width, height
        self._owner = _handle is None
        self._h = C.c_void_p()
        if _handle is None:
            capi.call("vkpbrt_illumination_buffer_create", ctx.handle, self.TYPE, width, height, C.byref(self._h))
        else:
            self._h = _handle
        n = C.c_uint32()
        t = C.c_uint32()
        capi.call("vkpbrt_illumination_buffer_type", self._h, C.byref(t), C.byref(n))
        self.type = t.value
        self.illumination_images: List[DescriptorImage] = [
            _borrow(ctx, "vkpbrt_illumination_buffer_image", self._h, i) for i in range(n.value)]

    @classmethod
    def create(cls, ctx: Context, width: int, height: int):
        return cls(ctx, width, height)

    @classmethod
    def from_images(cls, ctx: Context, images: List[DescriptorImage]):
        """wraps caller-provided images (imported / externally allocated memory)"""
        arr = (C.c_void_p * len(images))(*[i.handle for i in images])
        h = C.c_void_p()
        capi.call("vkpbrt_illumination_buffer_create_from_images", ctx.handle, cls.TYPE, arr, len(images), C.byref(h))
        i0 = images[0].info()
        self = cls(ctx, i0.width, i0.height, _handle=h)
        self._owner = True
        self._keep = list(images)
        return self

    @property
    def handle(self):
        return self._h

    def compile(self, context: Optional[Context] = None) -> None:
        capi.call("vkpbrt_illumination_buffer_compile", self._h)

    def update_image_layouts(self, context: Optional[Context] = None) -> None:
        pass

    def __del__(self):
        try:
            if self._owner:
                capi.lib().vkpbrt_illumination_buffer_destroy(self._h)
        except Exception:
            pass


class IlluminationBufferFinal(IlluminationBuffer):
    TYPE = capi.ILLUMINATION_FINAL


class IlluminationBufferFinalDemodulated(IlluminationBuffer):
    TYPE = capi.ILLUMINATION_FINAL_DEMODULATED


class IlluminationBufferDemodulated(IlluminationBuffer):
    """2 x rgba16f: illumination, illuminationSquared (IlluminationBuffer.cpp:223-258)."""
    TYPE = capi.ILLUMINATION_DEMODULATED


class IlluminationBufferDemodulatedFloat(IlluminationBuffer):
    """1 x rgba32f raw 1-spp demodulated illumination (IlluminationBuffer.cpp:260-282)."""
    TYPE = capi.ILLUMINATION_DEMODULATED_FLOAT


class Commands:
    """vsg::Commands: modules append their dispatches once; record() replays them every frame
    (viewer->recordAndSubmit(), source/VulkanPBRT.cpp:588).  Also carries the last bound push
    constants, which is how taa.comp gets its frameNumber in the reference."""

    def __init__(self):
        self.children: List[Callable[["Commands"], None]] = []
        self.bound_push_constants: Optional["PushConstants"] = None

    @classmethod
    def create(cls) -> "Commands":
        return cls()

    def add_child(self, fn: Callable[["Commands"], None]) -> None:
        self.children.append(fn)

    addChild = add_child

    def record(self) -> None:
        for c in self.children:
            c(self)


class AccumulationBuffer:
    """source/buffers/AccumulationBuffer.hpp:13-24."""

    def __init__(self, ctx: Context, width: int, height: int, _handle=None):
        self.ctx, self.width, self.height = ctx, width, height
        self._owner = _handle is None
        self._h = C.c_void_p()
        if _handle is None:
            capi.call("vkpbrt_accumulation_buffer_create", ctx.handle, width, height, C.byref(self._h))
        else:
            self._h = _handle
        g = lambda m: _borrow(ctx, "vkpbrt_accumulation_buffer_image", self._h, m)
        self.prev_illu = g(capi.ACC_PREV_ILLU)
        self.prev_illu_squared = g(capi.ACC_PREV_ILLU_SQUARED)
        self.prev_depth = g(capi.ACC_PREV_DEPTH)
        self.prev_normal = g(capi.ACC_PREV_NORMAL)
        self.spp = g(capi.ACC_SPP)
        self.prev_spp = g(capi.ACC_PREV_SPP)
        self.motion = g(capi.ACC_MOTION)
        self.next_depth = g(capi.ACC_NEXT_DEPTH)     # extension: depth history being written this frame

    @classmethod
    def create(cls, ctx: Context, width: int, height: int):
        return cls(ctx, width, height)

    @property
    def handle(self):
        return self._h

    def compile(self, context: Optional[Context] = None) -> None:
        capi.call("vkpbrt_accumulation_buffer_compile", self._h)

    def update_image_layouts(self, context: Optional[Context] = None) -> None:
        pass

    def copy_to_back_images(self, commands: Commands, g_buffer: GBuffer, illumination_buffer: IlluminationBuffer) -> None:
        """AccumulationBuffer.cpp:72-244: appended once at the end of the command list."""
        commands.add_child(lambda _c: capi.call("vkpbrt_accumulation_buffer_copy_to_back_images", self._h,
                                                g_buffer.handle, illumination_buffer.handle))

    def __del__(self):
        try:
            if self._owner:
                capi.lib().vkpbrt_accumulation_buffer_destroy(self._h)
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------
# per-frame constants
# ---------------------------------------------------------------------------------------------------
class PushConstants:
    """vsg::PushConstants holding RayTracingPushConstants (PipelineStructs.hpp:6-13); the frame loop
    mutates .value in place (VulkanPBRT.cpp:561-563)."""

    def __init__(self):
        self.value = _CPushConstants()

    @classmethod
    def create(cls):
        return cls()


class CameraMatrices:
    """source/io/RenderIO.hpp:28-34; column-major 16-float sequences, proj/inv_proj optional."""

    def __init__(self, view=None, inv_view=None, proj=None, inv_proj=None):
        self.view, self.inv_view, self.proj, self.inv_proj = view, inv_view, proj, inv_proj

    def to_c(self) -> _CCameraMatrices:
        c = _CCameraMatrices()
        z = capi._ZERO16
        c.view = capi.mat16(self.view if self.view is not None else z)
        c.inv_view = capi.mat16(self.inv_view if self.inv_view is not None else z)
        c.has_proj = 1 if (self.proj is not None and self.inv_proj is not None) else 0
        c.proj = capi.mat16(self.proj if self.proj is not None else z)
        c.inv_proj = capi.mat16(self.inv_proj if self.inv_proj is not None else z)
        return c


class DenoisingType(Enum):   # PipelineStructs.hpp:15-21
    NONE = 0
    BMFR = 1
    BFR = 2
    SVG = 3


class DenoisingBlockSize(Enum):   # PipelineStructs.hpp:23-29
    X8 = 0
    X16 = 1
    X32 = 2
    X64 = 3
    X8X16X32 = 4


# ---------------------------------------------------------------------------------------------------
# render modules
# ---------------------------------------------------------------------------------------------------
class Accumulator:
    """source/renderModules/Accumulator.hpp:15-25."""

    def __init__(self, g_buffer: GBuffer, illumination_buffer: IlluminationBuffer, separate_matrices: bool,
                 work_width: int = 16, work_height: int = 16):
        ctx = g_buffer.ctx
        self.ctx = ctx
        self._g, self._illum = g_buffer, illumination_buffer   # keep the borrowed buffers alive
        self._h = C.c_void_p()
        capi.call("vkpbrt_accumulator_create", ctx.handle, g_buffer.handle, illumination_buffer.handle,
                  1 if separate_matrices else 0, work_width, work_height, C.byref(self._h))
        hi = C.c_void_p()
        capi.call("vkpbrt_accumulator_accumulated_illumination", self._h, C.byref(hi))
        self.accumulated_illumination = IlluminationBufferDemodulated(ctx, g_buffer.width, g_buffer.height, _handle=hi)
        ha = C.c_void_p()
        capi.call("vkpbrt_accumulator_accumulation_buffer", self._h, C.byref(ha))
        self.accumulation_buffer = AccumulationBuffer(ctx, g_buffer.width, g_buffer.height, _handle=ha)
        # the bundles live inside the accumulator handle: whoever holds them keeps it alive
        self.accumulated_illumination._parent = self
        self.accumulation_buffer._parent = self

    @classmethod
    def create(cls, *a, **k):
        return cls(*a, **k)

    def compile_images(self, context: Optional[Context] = None) -> None:
        capi.call("vkpbrt_accumulator_compile_images", self._h)

    def update_image_layouts(self, context: Optional[Context] = None) -> None:
        pass

    def add_dispatch_to_command_graph(self, command_graph: Commands) -> None:
        command_graph.add_child(lambda _c: capi.call("vkpbrt_accumulator_record", self._h))

    def set_camera_matrices(self, frame_index: int, cur: CameraMatrices, prev: CameraMatrices) -> None:
        c, p = cur.to_c(), prev.to_c()
        capi.call("vkpbrt_accumulator_set_camera_matrices", self._h, int(frame_index), C.byref(c), C.byref(p))

    def set_row_range(self, row_begin: int, row_end: int) -> None:
        capi.call("vkpbrt_accumulator_set_row_range", self._h, row_begin, row_end)

    def set_force_scalar(self, enable: bool) -> None:
        """debug / test switch: the one-pixel-per-thread kernel with the IEEE library routines"""
        capi.call("vkpbrt_accumulator_set_force_scalar", self._h, 1 if enable else 0)

    def set_max_displacement_rows(self, rows: int) -> None:
        """band-sharded runs: count reprojection taps that leave the rows this rank holds (0 = off)"""
        capi.call("vkpbrt_accumulator_set_max_displacement_rows", self._h, int(rows))

    def displacement_violations(self) -> int:
        n = C.c_uint32(0)
        capi.call("vkpbrt_accumulator_displacement_violations", self._h, C.byref(n))
        return int(n.value)

    def __del__(self):
        try:
            capi.lib().vkpbrt_accumulator_destroy(self._h)   # also frees the two bundles it owns
        except Exception:
            pass


class _BlockDenoiser:
    _prefix = ""

    def compile(self, context: Optional[Context] = None) -> None:
        capi.call(f"vkpbrt_{self._prefix}_compile", self._h)

    def update_image_layouts(self, context: Optional[Context] = None) -> None:
        pass

    def add_dispatch_to_command_graph(self, command_graph: Commands, push_constants: PushConstants) -> None:
        def rec(c: Commands):
            c.bound_push_constants = push_constants
            capi.call(f"vkpbrt_{self._prefix}_record", self._h, C.byref(push_constants.value))
        command_graph.add_child(rec)

    def get_final_descriptor_image(self) -> DescriptorImage:
        return self._final

    def __del__(self):
        try:
            getattr(capi.lib(), f"vkpbrt_{self._prefix}_destroy")(self._h)
        except Exception:
            pass


class BMFR(_BlockDenoiser):
    """source/renderModules/denoisers/BMFR.hpp:17-25.  pre + fit + post run as one fused kernel."""
    _prefix = "bmfr"

    def __init__(self, width: int, height: int, work_width: int, work_height: int, g_buffer: GBuffer,
                 illu_buffer: IlluminationBuffer, acc_buffer: AccumulationBuffer, fitting_kernel: int = 256,
                 debug_outputs: int = 0):
        """debug_outputs: bit 0 materialises the reference's feature buffer / weights images, bit 1 routes every block
        through the out-of-line IEEE-division fit (qr_generic in bmfr.cu) so tests can cover it"""
        self.ctx = g_buffer.ctx
        self._keep = (g_buffer, illu_buffer, acc_buffer)
        self._h = C.c_void_p()
        capi.call("vkpbrt_bmfr_create", self.ctx.handle, width, height, work_width, work_height, g_buffer.handle,
                  illu_buffer.handle, acc_buffer.handle, fitting_kernel, C.byref(self._h))
        if debug_outputs:
            capi.call("vkpbrt_bmfr_set_debug_outputs", self._h, int(debug_outputs))
        self.kernel_name = f"k_bmfr_block<{work_width},{fitting_kernel}>"
        self._final = _borrow(self.ctx, "vkpbrt_bmfr_final_image", self._h)

    @classmethod
    def create(cls, *a, **k):
        return cls(*a, **k)

    def set_lane(self, lane: int) -> None:
        """0: the context's stream; 1, 2: a side lane that runs concurrently with the other denoisers of the frame"""
        capi.call("vkpbrt_bmfr_set_lane", self._h, int(lane))

    def set_position_type(self, position_type: int) -> None:
        """the shaders' POSITION_TYPE: 0 depth (default), 1 world with normalised depth, 2 world (bmfrGeneral.comp:30-31)"""
        capi.call("vkpbrt_bmfr_set_position_type", self._h, int(position_type))

    def set_block_row_range(self, begin: int, end: int) -> None:
        capi.call("vkpbrt_bmfr_set_block_row_range", self._h, begin, end)

    def image(self, which: int) -> DescriptorImage:
        return _borrow(self.ctx, "vkpbrt_bmfr_image_get", self._h, which)

    @property
    def denoised(self) -> DescriptorImage:
        return self.image(capi.BMFR_IMAGE_DENOISED)

    @property
    def feature_buffer(self) -> DescriptorImage:
        return self.image(capi.BMFR_IMAGE_FEATURES)

    @property
    def weights(self) -> DescriptorImage:
        return self.image(capi.BMFR_IMAGE_WEIGHTS)


class BFR(_BlockDenoiser):
    """source/renderModules/denoisers/BFR.hpp:11-18."""
    _prefix = "bfr"

    def __init__(self, width: int, height: int, work_width: int, work_height: int, g_buffer: GBuffer,
                 illu_buffer: IlluminationBuffer, acc_buffer: AccumulationBuffer):
        self.ctx = g_buffer.ctx
        self._keep = (g_buffer, illu_buffer, acc_buffer)
        self._h = C.c_void_p()
        capi.call("vkpbrt_bfr_create", self.ctx.handle, width, height, work_width, work_height, g_buffer.handle,
                  illu_buffer.handle, acc_buffer.handle, C.byref(self._h))
        self.kernel_name = f"k_bfr_block<{work_width}>"
        self._final = _borrow(self.ctx, "vkpbrt_bfr_final_image", self._h)

    @classmethod
    def create(cls, *a, **k):
        return cls(*a, **k)

    def set_lane(self, lane: int) -> None:
        """0: the context's stream; 1, 2: a side lane that runs concurrently with the other denoisers of the frame"""
        capi.call("vkpbrt_bfr_set_lane", self._h, int(lane))

    @property
    def denoised(self) -> DescriptorImage:
        return _borrow(self.ctx, "vkpbrt_bfr_denoised_image", self._h)


class BFRBlender:
    """source/renderModules/denoisers/BFRBlender.hpp:9-18."""

    def __init__(self, width: int, height: int, average_image: DescriptorImage, average_squared_image: DescriptorImage,
                 denoised0: DescriptorImage, denoised1: DescriptorImage, denoised2: DescriptorImage,
                 work_width: int = 16, work_height: int = 16, filter_radius: int = 2):
        self.ctx = average_image.ctx
        self._keep = (average_image, average_squared_image, denoised0, denoised1, denoised2)
        self._h = C.c_void_p()
        capi.call("vkpbrt_bfr_blender_create", self.ctx.handle, width, height, average_image.handle,
                  average_squared_image.handle, denoised0.handle, denoised1.handle, denoised2.handle, work_width,
                  work_height, filter_radius, C.byref(self._h))
        self.kernel_name = "k_bfr_blend"
        self._final = _borrow(self.ctx, "vkpbrt_bfr_blender_final_image", self._h)

    @classmethod
    def create(cls, *a, **k):
        return cls(*a, **k)

    def compile(self, context: Optional[Context] = None) -> None:
        capi.call("vkpbrt_bfr_blender_compile", self._h)

    def update_image_layouts(self, context: Optional[Context] = None) -> None:
        pass

    def add_dispatch_to_command_graph(self, command_graph: Commands) -> None:
        command_graph.add_child(lambda _c: capi.call("vkpbrt_bfr_blender_record", self._h))

    def copy_final_image(self, commands: Commands, dst_image: DescriptorImage) -> None:
        """BFRBlender.hpp:17 / BFRBlender.cpp:91-124: appends a copy of the final image into ``dst_image``."""
        src = self._final
        commands.add_child(lambda _c: capi.call("vkpbrt_image_copy_record", src.handle, dst_image.handle))

    def get_final_descriptor_image(self) -> DescriptorImage:
        return self._final

    def __del__(self):
        try:
            capi.lib().vkpbrt_bfr_blender_destroy(self._h)
        except Exception:
            pass


class Taa:
    """source/renderModules/Taa.hpp:14-20."""

    def __init__(self, width: int, height: int, work_width: int, work_height: int, g_buffer: GBuffer,
                 acc_buffer: AccumulationBuffer, denoised: DescriptorImage, fix_swizzle: bool = False):
        self.ctx = g_buffer.ctx
        self._keep = (g_buffer, acc_buffer, denoised)
        self._h = C.c_void_p()
        capi.call("vkpbrt_taa_create", self.ctx.handle, width, height, work_width, work_height, g_buffer.handle,
                  acc_buffer.handle, denoised.handle, C.byref(self._h))
        if fix_swizzle:
            capi.call("vkpbrt_taa_set_fix_swizzle", self._h, 1)
        self._final = _borrow(self.ctx, "vkpbrt_taa_final_image", self._h)
        self.history = _borrow(self.ctx, "vkpbrt_taa_history_image", self._h)

    @classmethod
    def create(cls, *a, **k):
        return cls(*a, **k)

    def compile(self, context: Optional[Context] = None) -> None:
        capi.call("vkpbrt_taa_compile", self._h)

    def update_image_layouts(self, context: Optional[Context] = None) -> None:
        pass

    def set_row_range(self, row_begin: int, row_end: int) -> None:
        capi.call("vkpbrt_taa_set_row_range", self._h, row_begin, row_end)

    def set_force_scalar(self, enable: bool) -> None:
        """debug / test switch: the one-pixel-per-thread kernel"""
        capi.call("vkpbrt_taa_set_force_scalar", self._h, 1 if enable else 0)

    def set_strip_rows(self, rows: int) -> None:
        """test switch: rows per warp and launch of the two-column kernel (0 = automatic); results do not depend on it"""
        capi.call("vkpbrt_taa_set_strip_rows", self._h, int(rows))

    def record_part(self, push_constants: "PushConstants", row_begin: int, row_end: int, last: bool) -> None:
        """one of several launches of a frame's TAA over disjoint row ranges (see vkpbrt_taa_record_part)"""
        capi.call("vkpbrt_taa_record_part", self._h, C.byref(push_constants.value), int(row_begin), int(row_end), 1 if last else 0)

    def record_parts(self, push_constants: "PushConstants", row_begin: int, row_end: int, row_begin2: int, row_end2: int, last: bool) -> None:
        """two disjoint row ranges in ONE launch (either may be empty): the first and the last row of a band"""
        capi.call("vkpbrt_taa_record_parts", self._h, C.byref(push_constants.value), int(row_begin), int(row_end), int(row_begin2),
                  int(row_end2), 1 if last else 0)

    def add_dispatch_to_command_graph(self, command_graph: Commands) -> None:
        def rec(c: Commands):
            pc = c.bound_push_constants
            if pc is None:
                raise VkpbrtError(capi.ERR_INVALID_ARGUMENT,
                                  "Taa: no push constants bound; record a denoiser before Taa (Taa.cpp:99-107)")
            capi.call("vkpbrt_taa_record", self._h, C.byref(pc.value))
        command_graph.add_child(rec)

    def copy_final_image(self, commands: Commands, dst_image: DescriptorImage) -> None:
        """Taa.hpp:23 / Taa.cpp:108-141: appends a copy of the final image into ``dst_image`` (the reference uses it for
        its own final -> history copy, which the ping-pong pair replaces here, and exposes it publicly)."""
        src = self._final
        commands.add_child(lambda _c: capi.call("vkpbrt_image_copy_record", src.handle, dst_image.handle))

    def get_final_descriptor_image(self) -> DescriptorImage:
        return self._final

    def __del__(self):
        try:
            capi.lib().vkpbrt_taa_destroy(self._h)
        except Exception:
            pass


class FormatConverter:
    """source/renderModules/FormatConverter.hpp:4-21: converts the final image to B8G8R8A8_UNORM when no denoiser produced
    one (VulkanPBRT.cpp:476-484)."""

    def __init__(self, src_image: DescriptorImage, dst_format: int = capi.FORMAT_B8G8R8A8_UNORM, work_width: int = 16,
                 work_height: int = 16):
        self.ctx = src_image.ctx
        self._keep = src_image
        self._h = C.c_void_p()
        capi.call("vkpbrt_format_converter_create", self.ctx.handle, src_image.handle, dst_format, work_width, work_height,
                  C.byref(self._h))
        self.final_image = _borrow(self.ctx, "vkpbrt_format_converter_final_image", self._h)

    @classmethod
    def create(cls, *a, **k):
        return cls(*a, **k)

    def compile_images(self, context: Optional[Context] = None) -> None:
        capi.call("vkpbrt_format_converter_compile_images", self._h)

    def update_image_layouts(self, context: Optional[Context] = None) -> None:
        pass

    def add_dispatch_to_command_graph(self, command_graph: Commands) -> None:
        command_graph.add_child(lambda _c: capi.call("vkpbrt_format_converter_record", self._h))

    def __del__(self):
        try:
            capi.lib().vkpbrt_format_converter_destroy(self._h)
        except Exception:
            pass


def demodulate(ctx: Context, radiance: DescriptorImage, albedo: DescriptorImage, position_x: DescriptorImage,
               demodulated: DescriptorImage) -> None:
    """the producer-side convention of shaders/ptRaygen.rgen:81-88 (see vkpbrt_demodulate_record)"""
    capi.call("vkpbrt_demodulate_record", ctx.handle, radiance.handle, albedo.handle, position_x.handle, demodulated.handle)


# ---------------------------------------------------------------------------------------------------
# source/util/DenoiserUtils.cpp:8-130
# ---------------------------------------------------------------------------------------------------
def add_denoiser_to_commands(denoising_type: DenoisingType, denoising_size: DenoisingBlockSize, commands: Commands,
                             compile_context: Context, width: int, height: int, compute_constants: PushConstants,
                             g_buffer: GBuffer, illumination_buffer: IlluminationBuffer,
                             accumulation_buffer: AccumulationBuffer, average_squared_image: Optional[DescriptorImage] = None):
    """vkpbrt::add_denoiser_to_commands.  Returns (final_descriptor_image, modules) -- the reference
    returns the final image through an out-parameter and leaks the modules into the command graph;
    the caller must keep `modules` alive here.

    X8X16X32 blends three block sizes with BFRBlender.  The reference feeds the blender
    illumination_images[1] ("illuminationSquared"), which no shader ever writes;
    pass average_squared_image to supply a defined second-moment plane instead."""
    sizes = {DenoisingBlockSize.X8: 8, DenoisingBlockSize.X16: 16, DenoisingBlockSize.X32: 32}
    if denoising_type == DenoisingType.NONE:
        return None, []
    if denoising_type == DenoisingType.SVG:
        print("Not yet implemented")   # DenoiserUtils.cpp:126
        return None, []
    if denoising_type == DenoisingType.BFR:
        make = lambda b: BFR.create(width, height, b, b, g_buffer, illumination_buffer, accumulation_buffer)
    else:
        make = lambda b: BMFR.create(width, height, b, b, g_buffer, illumination_buffer, accumulation_buffer, 64 if b == 8 else 256)
    if denoising_size in sizes:
        d = make(sizes[denoising_size])
        d.compile(compile_context)
        d.update_image_layouts(compile_context)
        d.add_dispatch_to_command_graph(commands, compute_constants)
        return d.get_final_descriptor_image(), [d]
    if denoising_size == DenoisingBlockSize.X8X16X32:
        d8, d16, d32 = make(8), make(16), make(32)
        # the three block sizes are independent: b = 16 and b = 32 run on side lanes, concurrently with b = 8
        d16.set_lane(1)
        d32.set_lane(2)
        avg = illumination_buffer.illumination_images[0]
        avg_sq = average_squared_image if average_squared_image is not None else illumination_buffer.illumination_images[1]
        blender = BFRBlender.create(width, height, avg, avg_sq, d8.get_final_descriptor_image(),
                                    d16.get_final_descriptor_image(), d32.get_final_descriptor_image())
        for m in (d8, d16, d32, blender):
            m.compile(compile_context)
            m.update_image_layouts(compile_context)
        for m in (d8, d16, d32):
            m.add_dispatch_to_command_graph(commands, compute_constants)
        blender.add_dispatch_to_command_graph(commands)
        return blender.get_final_descriptor_image(), [d8, d16, d32, blender]
    raise VkpbrtError(capi.ERR_UNSUPPORTED, f"unsupported denoising block size {denoising_size}")
