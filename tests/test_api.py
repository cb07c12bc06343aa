"""C-ABI surface and host-logic tests: export check (CPU, no compute), error behaviour mirroring the
reference's constructors, image plumbing, band (row-range) dispatch."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from tests.conftest import backend_params

ROOT = Path(__file__).resolve().parents[1]


def test_c_abi_library_exports_every_declared_symbol():
    """no GPU needed: the product library loads and exports exactly what include/vkpbrt_b200.h declares"""
    header = (ROOT / "include" / "vkpbrt_b200.h").read_text()
    declared = set(re.findall(r"VKPBRT_API\s+[\w\s\*]+?\b(vkpbrt_\w+)\s*\(", header))
    assert len(declared) > 60
    lib = ctypes.CDLL(str(ROOT / "vulkanpbrt_b200" / "lib" / "libvkpbrt_b200.so"))
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    from vulkanpbrt_b200 import _capi
    assert set(_capi.EXPORTS) == declared, set(_capi.EXPORTS) ^ declared


def test_product_library_has_no_cpu_fallback():
    """without a CUDA device the context cannot be created: the package fails loudly"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vulkanpbrt_b200 import Context, VkpbrtError, _capi
    assert _capi._lib is None or b"HOSTSIM" not in _capi._lib.vkpbrt_version()
    with pytest.raises(VkpbrtError) as e:
        Context(0)
    assert e.value.code == _capi.ERR_NO_DEVICE


def test_halo_entry_points_reject_null_arguments_without_a_device():
    """argument checks of the multi-GPU entry points come before any CUDA call: callable on a machine without a GPU"""
    from vulkanpbrt_b200 import _capi
    lib = _capi.configure(ctypes.CDLL(str(ROOT / "vulkanpbrt_b200" / "lib" / "libvkpbrt_b200.so")))
    out = ctypes.c_void_p()
    off = ctypes.c_uint64()
    buf = (ctypes.c_uint8 * 64)()
    assert lib.vkpbrt_peer_export(None, None, buf, ctypes.byref(off)) == _capi.ERR_INVALID_ARGUMENT
    assert lib.vkpbrt_peer_open(None, buf, ctypes.byref(out)) == _capi.ERR_INVALID_ARGUMENT
    assert lib.vkpbrt_halo_exchange_create(None, None, 1000, ctypes.byref(out)) == _capi.ERR_INVALID_ARGUMENT
    assert lib.vkpbrt_halo_exchange_start(None, None, None, 1) == _capi.ERR_INVALID_ARGUMENT
    assert lib.vkpbrt_halo_exchange_wait(None, None, 1) == _capi.ERR_INVALID_ARGUMENT
    assert lib.vkpbrt_halo_exchange_destroy(None) == _capi.OK
    assert b"null" in lib.vkpbrt_last_error()


def test_product_library_contains_sm100a_code_only():
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-lelf", str(ROOT / "vulkanpbrt_b200" / "lib" / "libvkpbrt_b200.so")],
                         capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
class TestHostBehaviour:
    def test_image_upload_download_roundtrip(self, backend):
        from vulkanpbrt_b200 import Context, DescriptorImage, _capi
        ctx = Context(0)
        img = DescriptorImage.create(ctx, _capi.FORMAT_R16G16B16A16_SFLOAT, 37, 19, layers=2)
        img.compile()
        assert (img.download() == 0).all()            # compile() zero-initialises (documented initial history)
        a = np.arange(2 * 19 * 37 * 4, dtype=np.uint16).reshape(2, 19, 37, 4)
        img.upload(a)
        np.testing.assert_array_equal(img.download(), a)
        i = img.info()
        assert (i.width, i.height, i.layers, i.row_pitch, i.size_bytes) == (37, 19, 2, 37 * 8, 2 * 19 * 37 * 8)

    def test_copy_final_image_and_device_uuid(self, backend):
        """Taa.hpp:23 / BFRBlender.hpp:17 copy_final_image: a recorded whole-image copy between size-compatible formats;
        vkpbrt_device_uuid: what a Vulkan host matches VkPhysicalDeviceIDProperties::deviceUUID against"""
        from vulkanpbrt_b200 import Commands, Context, DenoisePipeline, DenoisingType, DescriptorImage, VkpbrtError, _capi, synth
        n = ctypes.c_int(0)
        _capi.call("vkpbrt_device_count", ctypes.byref(n))
        assert n.value >= 1
        uuid = (ctypes.c_uint8 * 16)()
        _capi.call("vkpbrt_device_uuid", 0, uuid)
        assert any(uuid)
        assert _capi.lib().vkpbrt_device_uuid(n.value, uuid) == _capi.ERR_INVALID_ARGUMENT
        W, H = 192, 128
        pipe = DenoisePipeline(W, H, DenoisingType.BMFR, use_taa=True)
        dst = DescriptorImage.create(pipe.ctx, _capi.FORMAT_R8G8B8A8_UNORM, W, H)       # 4-byte texels like the BGRA8 final
        dst.compile()
        pipe.taa.copy_final_image(pipe.commands, dst)
        for f in range(2):
            pipe.run_frame(f, synth.render_frame(W, H, f))
            pipe.ctx.synchronize()
            np.testing.assert_array_equal(dst.download(), pipe.final.download())
        small = DescriptorImage.create(pipe.ctx, _capi.FORMAT_R8G8B8A8_UNORM, W, H - 1)
        small.compile()
        with pytest.raises(VkpbrtError) as e:
            _capi.call("vkpbrt_image_copy_record", pipe.final.handle, small.handle)
        assert e.value.code == _capi.ERR_INVALID_ARGUMENT

    def test_wrong_illumination_buffer_type_is_rejected(self, backend):
        """denoisers/BMFR.cpp:17-22 / BFR.cpp:15-20 print and return a half-built object; we reject"""
        from vulkanpbrt_b200 import (BFR, BMFR, AccumulationBuffer, Context, GBuffer, IlluminationBufferFinal, VkpbrtError,
                                     _capi)
        ctx = Context(0)
        g, acc, ill = GBuffer.create(ctx, 64, 64), AccumulationBuffer.create(ctx, 64, 64), IlluminationBufferFinal.create(ctx, 64, 64)
        for cls in (BMFR, BFR):
            with pytest.raises(VkpbrtError) as e:
                cls.create(64, 64, 32, 32, g, ill, acc)
            assert e.value.code == _capi.ERR_WRONG_BUFFER_TYPE
            assert "IlluminationBufferDemodulated" in str(e.value)

    def test_unsupported_block_size_and_record_before_compile(self, backend):
        from vulkanpbrt_b200 import (BMFR, Accumulator, Commands, Context, GBuffer, IlluminationBufferDemodulatedFloat,
                                     PushConstants, VkpbrtError, _capi)
        ctx = Context(0)
        g, raw = GBuffer.create(ctx, 64, 64), IlluminationBufferDemodulatedFloat.create(ctx, 64, 64)
        acc = Accumulator.create(g, raw, True)
        with pytest.raises(VkpbrtError) as e:
            BMFR.create(64, 64, 64, 64, g, acc.accumulated_illumination, acc.accumulation_buffer)
        assert e.value.code == _capi.ERR_UNSUPPORTED
        bmfr = BMFR.create(64, 64, 32, 32, g, acc.accumulated_illumination, acc.accumulation_buffer)
        cmds = Commands.create()
        bmfr.add_dispatch_to_command_graph(cmds, PushConstants.create())
        with pytest.raises(VkpbrtError) as e:
            cmds.record()
        assert e.value.code == _capi.ERR_NOT_COMPILED
        cmds2 = Commands.create()
        acc.add_dispatch_to_command_graph(cmds2)
        with pytest.raises(VkpbrtError) as e:
            cmds2.record()
        assert e.value.code == _capi.ERR_NOT_COMPILED

    def test_missing_separate_matrices_raises(self, backend):
        """Accumulator.cpp:89-94 throws when created with separate_matrices but fed combined matrices"""
        from vulkanpbrt_b200 import (Accumulator, CameraMatrices, Context, GBuffer, IlluminationBufferDemodulatedFloat,
                                     VkpbrtError, _capi)
        ctx = Context(0)
        g, raw = GBuffer.create(ctx, 32, 32), IlluminationBufferDemodulatedFloat.create(ctx, 32, 32)
        acc = Accumulator.create(g, raw, True)
        eye = list(np.eye(4, dtype=np.float32).reshape(-1))
        with pytest.raises(VkpbrtError) as e:
            acc.set_camera_matrices(0, CameraMatrices(view=eye, inv_view=eye), CameraMatrices(view=eye, inv_view=eye))
        assert e.value.code == _capi.ERR_MISSING_MATRICES

    def test_taa_without_a_denoiser_has_no_push_constants(self, backend):
        """Taa.cpp:99-107 never pushes constants itself"""
        from vulkanpbrt_b200 import DenoisePipeline, DenoisingType, Commands, Taa, VkpbrtError
        pipe = DenoisePipeline(64, 64, DenoisingType.BMFR, use_taa=False)
        taa = Taa.create(64, 64, 16, 16, pipe.g_buffer, pipe.accumulation_buffer, pipe.final)
        taa.compile()
        cmds = Commands.create()
        taa.add_dispatch_to_command_graph(cmds)
        with pytest.raises(VkpbrtError):
            cmds.record()

    def test_band_dispatch_equals_full_frame(self, backend, oracle):
        """row / block-row ranges (multi-GPU band sharding) are a pure partition of the launch grid"""
        from vulkanpbrt_b200 import DenoisePipeline, synth
        W, H = 160, 192
        full = DenoisePipeline(W, H, use_taa=True)
        band = DenoisePipeline(W, H, use_taa=True)
        nby = H // 32 + 2
        for f in range(3):
            fr = synth.render_frame(W, H, f)
            full.run_frame(f, fr)
            band.upload_frame(fr)
            band.set_frame_constants(f, fr.camera)
            # two bands, each module recorded band by band (all of accumulate first: BMFR reads across bands)
            for r0, r1 in ((0, 70), (70, H)):
                band.accumulator.set_row_range(r0, r1)
                band.commands.children[0](band.commands)
            for b0, b1 in ((0, 3), (3, nby)):
                band.modules[0].set_block_row_range(b0, b1)
                band.commands.children[1](band.commands)
            for r0, r1 in ((0, 101), (101, H)):
                band.taa.set_row_range(r0, r1)
                band.commands.children[2](band.commands)
                # both bands of one frame must read the same history: undo the ping-pong flip of the first
                if r1 != H:
                    first_out = band.taa.get_final_descriptor_image().device_ptr
            band.commands.children[3](band.commands)
            band.end_frame(fr.camera)
            full.ctx.synchronize(); band.ctx.synchronize()
            np.testing.assert_array_equal(band.denoiser_final.download(), full.denoiser_final.download())
            np.testing.assert_array_equal(band.modules[0].denoised.download(), full.modules[0].denoised.download())
            np.testing.assert_array_equal(band.accumulation_buffer.motion.download(), full.accumulation_buffer.motion.download())


def test_destroying_the_modules_frees_every_device_allocation():
    """emulator-only: the emulator counts device allocations; after the pipelines of every wiring (incl. the debug images,
    the per-frame tables, the side lanes) are destroyed none is left"""
    import gc
    import subprocess
    from tests.conftest import HOSTSIM_DIR, HOSTSIM_LIB
    from vulkanpbrt_b200 import DenoisePipeline, DenoisingBlockSize, DenoisingType, _capi, synth
    subprocess.run(["make", "-C", str(HOSTSIM_DIR)], check=True, capture_output=True)
    gc.collect()
    saved = _capi._lib
    lib = ctypes.CDLL(str(HOSTSIM_LIB))
    lib.hostsim_live_allocations.restype = ctypes.c_long
    _capi._lib = _capi.configure(lib)
    try:
        before = lib.hostsim_live_allocations()
        for den, bs, taa in [(DenoisingType.BMFR, DenoisingBlockSize.X32, True), (DenoisingType.BFR, DenoisingBlockSize.X16, False),
                             (DenoisingType.BFR, DenoisingBlockSize.X8X16X32, True), (DenoisingType.BMFR, DenoisingBlockSize.X8X16X32, True)]:
            x3 = bs == DenoisingBlockSize.X8X16X32
            pipe = DenoisePipeline(64, 64, den, bs, use_taa=taa, average_squared=x3, bmfr_debug_outputs=True)
            for f in range(2):
                pipe.run_frame(f, synth.render_frame(64, 64, f))
            pipe.ctx.synchronize()
            assert lib.hostsim_live_allocations() > before
            del pipe
            gc.collect()
            assert lib.hostsim_live_allocations() == before
    finally:
        gc.collect()
        _capi._lib = saved
