"""DenoisePipeline: plays the role of the reference's main() for the denoising path -- the setup
call order of source/VulkanPBRT.cpp:337-505 and the per-frame loop of :551-618 (live mode,
SEPARATE_MATRICES) -- on top of the module classes in vulkanpbrt_b200.modules.  Used by the tests,
__graft_entry__.smoke() and bench.py; applications embed the modules directly."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import _capi as capi
from .modules import (BFR, BMFR, Accumulator, BFRBlender, CameraMatrices, Commands, Context, DenoisingBlockSize,
                      DenoisingType, DescriptorImage, GBuffer, IlluminationBufferDemodulated,
                      IlluminationBufferDemodulatedFloat, PushConstants, Taa, add_denoiser_to_commands)

IDENTITY = [1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0]


class DenoisePipeline:
    def __init__(self, width: int, height: int, denoiser: DenoisingType = DenoisingType.BMFR,
                 block_size: DenoisingBlockSize = DenoisingBlockSize.X32, use_taa: bool = False, device: int = 0,
                 stream: Optional[int] = None, separate_matrices: bool = True, raw_f16: bool = False,
                 fix_taa_swizzle: bool = False, bmfr_debug_outputs: bool = False, average_squared: bool = False,
                 ctx: Optional[Context] = None, external_inputs: bool = False, position_type: int = 0):
        self.width, self.height = width, height
        self.ctx = ctx if ctx is not None else Context(device, stream)
        self.separate_matrices = separate_matrices
        ctx = self.ctx
        # VulkanPBRT.cpp:337-339 (live) / :375-388 (offline rgba16f input)
        raw_cls = IlluminationBufferDemodulated if raw_f16 else IlluminationBufferDemodulatedFloat
        if external_inputs:
            # the producer owns the input planes (Vulkan-imported memory, a resident sequence, ...): wrap them and
            # re-point the handles per frame with bind_inputs()
            F = capi
            wrap = lambda fmt: DescriptorImage.wrap(ctx, fmt, width, height, 0)
            self._ext = [wrap(F.FORMAT_R32_SFLOAT), wrap(F.FORMAT_R32G32_SFLOAT), wrap(F.FORMAT_R8G8B8A8_UNORM),
                         wrap(F.FORMAT_R16G16B16A16_SFLOAT if raw_f16 else F.FORMAT_R32G32B32A32_SFLOAT)]
            self.g_buffer = GBuffer.from_images(ctx, self._ext[0], self._ext[1], None, self._ext[2])
            self.raw_illumination = raw_cls.from_images(ctx, [self._ext[3]])
        else:
            self.g_buffer = GBuffer.create(ctx, width, height)
            self.raw_illumination = raw_cls.create(ctx, width, height)
            self.g_buffer.compile(ctx)
            self.raw_illumination.compile(ctx)
        self.commands = Commands.create()
        self.push_constants = PushConstants.create()
        self.push_constants.value.prev_view = capi.mat16(IDENTITY)
        # :426-432
        self.accumulator = Accumulator.create(self.g_buffer, self.raw_illumination, separate_matrices)
        self.accumulator.compile_images(ctx)
        self.accumulator.update_image_layouts(ctx)
        self.accumulator.add_dispatch_to_command_graph(self.commands)
        self.illumination_buffer = self.accumulator.accumulated_illumination
        self.accumulation_buffer = self.accumulator.accumulation_buffer
        # optional explicit second-moment plane for the blender
        self.average_squared_image: Optional[DescriptorImage] = None
        if average_squared:
            self.average_squared_image = DescriptorImage.create(ctx, capi.FORMAT_R16G16B16A16_SFLOAT, width, height)
            self.average_squared_image.compile()
        # :443
        self.modules: List = []
        self.final = None
        if denoiser == DenoisingType.BMFR and block_size != DenoisingBlockSize.X8X16X32 and bmfr_debug_outputs:
            b = {DenoisingBlockSize.X8: 8, DenoisingBlockSize.X16: 16, DenoisingBlockSize.X32: 32}[block_size]
            d = BMFR.create(width, height, b, b, self.g_buffer, self.illumination_buffer, self.accumulation_buffer,
                            64 if b == 8 else 256, debug_outputs=int(bmfr_debug_outputs))
            d.compile(ctx)
            d.add_dispatch_to_command_graph(self.commands, self.push_constants)
            self.final, self.modules = d.get_final_descriptor_image(), [d]
        else:
            self.final, self.modules = add_denoiser_to_commands(
                denoiser, block_size, self.commands, ctx, width, height, self.push_constants, self.g_buffer,
                self.illumination_buffer, self.accumulation_buffer, self.average_squared_image)
        if position_type:
            for m in self.modules:
                if isinstance(m, BMFR):
                    m.set_position_type(position_type)
        self.denoiser_final = self.final
        # :448-456
        self.taa: Optional[Taa] = None
        if use_taa and self.final is not None:
            self.taa = Taa.create(width, height, 16, 16, self.g_buffer, self.accumulation_buffer, self.final,
                                  fix_swizzle=fix_taa_swizzle)
            self.taa.compile(ctx)
            self.taa.update_image_layouts(ctx)
            self.taa.add_dispatch_to_command_graph(self.commands)
            self.final = self.taa.get_final_descriptor_image()
        # :502-505
        self.accumulation_buffer.copy_to_back_images(self.commands, self.g_buffer, self.illumination_buffer)
        # one label per recorded command, in replay order (bench.py's per-kernel timing)
        self.command_labels = (["k_accumulate"] + [m.kernel_name for m in self.modules] + (["k_taa"] if self.taa else [])
                               + ["copy_to_back_images (pointer swaps)"])
        assert len(self.command_labels) == len(self.commands.children)
        self._prev_camera = None
        ctx.synchronize()

    # ---- per frame: VulkanPBRT.cpp:551-618 ------------------------------------------------------------
    def set_frame_constants(self, frame_index: int, cam) -> None:
        pc = self.push_constants.value
        pc.view_inverse = capi.mat16(cam.inv_view)        # :561
        pc.proj_inverse = capi.mat16(cam.inv_proj)
        pc.frame_number = frame_index                     # :562
        pc.sample_number = 0
        prev = self._prev_camera
        if self.separate_matrices:
            a = CameraMatrices(inv_view=cam.inv_view, proj=cam.proj, inv_proj=cam.inv_proj)    # :578-583
            b = CameraMatrices(view=prev.view if prev is not None else IDENTITY)
        else:
            # offline mode (:572-573): combined view-projection matrices
            vp, ivp = _combined(cam)
            a = CameraMatrices(view=vp, inv_view=ivp)
            if prev is not None:
                pvp, pivp = _combined(prev)
                b = CameraMatrices(view=pvp, inv_view=pivp)
            else:
                b = a
        self.accumulator.set_camera_matrices(frame_index, a, b)

    def end_frame(self, cam) -> None:
        self.push_constants.value.prev_view = capi.mat16(cam.view)    # :591
        self._prev_camera = cam

    def run_frame_with_matrices(self, frame_index: int, frame, matrices) -> None:
        """offline mode as main() drives it (VulkanPBRT.cpp:566-573): the frame's planes are staged and the accumulator
        gets camera_matrices[f] and camera_matrices[f-1] straight from the imported file (matrix_io.import_matrices)"""
        self.upload_frame(frame)
        cur = matrices[frame_index]
        prev = matrices[frame_index - 1] if frame_index > 0 else cur
        pc = self.push_constants.value
        pc.view_inverse = capi.mat16(cur.inv_view)
        if cur.inv_proj is not None:
            pc.proj_inverse = capi.mat16(cur.inv_proj)
        pc.frame_number = frame_index
        pc.sample_number = 0
        self.accumulator.set_camera_matrices(frame_index, cur, prev)
        self.record()
        pc.prev_view = capi.mat16(cur.view)

    def upload_frame(self, frame) -> None:
        """stager->transfer_staging_data_from(frame) (:568-569)"""
        self.g_buffer.depth.upload(frame.depth, sync=False)
        self.g_buffer.normal.upload(frame.normal, sync=False)
        self.g_buffer.albedo.upload(frame.albedo, sync=False)
        img = self.raw_illumination.illumination_images[0]
        if img.info().format == capi.FORMAT_R16G16B16A16_SFLOAT:
            self._f16_tmp = frame.illumination.astype(np.float16)
            img.upload(self._f16_tmp, sync=False)
        else:
            img.upload(frame.illumination, sync=False)
        self.ctx.synchronize()

    def bind_inputs(self, depth_ptr: int, normal_ptr: int, albedo_ptr: int, illumination_ptr: int) -> None:
        """external_inputs mode: point the G-buffer / raw illumination handles at this frame's device planes"""
        for img, ptr in zip(self._ext, (depth_ptr, normal_ptr, albedo_ptr, illumination_ptr)):
            img.set_data(ptr)

    def record(self) -> None:
        self.commands.record()     # viewer->recordAndSubmit() (:588)

    def run_frame(self, frame_index: int, frame) -> None:
        self.upload_frame(frame)
        self.set_frame_constants(frame_index, frame.camera)
        self.record()
        self.end_frame(frame.camera)


def _combined(cam):
    """combined VP and its inverse for the accumulator's non-SEPARATE_MATRICES mode.  The shader's
    ray reconstruction (accumulator.comp:56-63) is calibrated to the BMFR data set's camera files;
    with vsg-style projections it needs the w row negated to reproduce the generator's rays
    (SURVEY.md App. A.0 option ii)."""
    v = np.asarray(cam.view, dtype=np.float64).reshape(4, 4).T
    p = np.asarray(cam.proj, dtype=np.float64).reshape(4, 4).T
    vp = p @ v
    ivp = np.linalg.inv(vp)
    return vp.T.astype(np.float32).reshape(-1), ivp.T.astype(np.float32).reshape(-1)
