"""FormatConverter (shaders/formatConverter.comp, source/renderModules/FormatConverter.cpp) and the producer-side
demodulation convention (shaders/ptRaygen.rgen:81-88): oracle pinned against the reference's shader source where the
shim can compile it, CUDA path bit-exact against the oracle."""
import numpy as np
import pytest

from tests.conftest import backend_params
from vulkanpbrt_b200 import Context, DescriptorImage, FormatConverter, VkpbrtError, _capi as capi, demodulate
from vulkanpbrt_b200.modules import Commands


def _test_image(H, W, seed=3):
    rng = np.random.default_rng(seed)
    a = rng.uniform(-0.25, 1.25, (H, W, 4)).astype(np.float32)
    a[0, 0] = [np.nan, np.inf, -np.inf, 0.5]          # NaN -> 0, clamps
    a[0, 1] = [0.0, 1.0, 0.5 / 255.0, 254.5 / 255.0]  # rounding boundaries
    a[1, :, :] = (np.arange(W)[:, None] / 255.0 + np.array([0, 1e-4, -1e-4, 0.5 / 255.0])[None, :]).astype(np.float32)
    return a


@pytest.mark.parametrize("dtype", ["f32", "f16", "u8"])
def test_format_converter_oracle_equals_reference_shader(oracle, dtype):
    from oracle import ref as R
    if not R.build():
        pytest.skip("oracle/_ref is not built and /root/reference is not mounted")
    a = _test_image(37, 50)
    src = a if dtype == "f32" else (a.astype(np.float16).view(np.uint16) if dtype == "f16" else (np.clip(np.nan_to_num(a), 0, 1) * 255).astype(np.uint8))
    np.testing.assert_array_equal(oracle.format_converter(src), R.format_converter(src))


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("dtype", ["f32", "f16", "u8"])
def test_format_converter_module(backend, oracle, dtype):
    """FormatConverter::create(src, VK_FORMAT_B8G8R8A8_UNORM) -> compile_images -> add_dispatch_to_command_graph
    (VulkanPBRT.cpp:476-484); odd size"""
    H, W = 37, 50
    a = _test_image(H, W)
    fmt, src = {"f32": (capi.FORMAT_R32G32B32A32_SFLOAT, a), "f16": (capi.FORMAT_R16G16B16A16_SFLOAT, a.astype(np.float16).view(np.uint16)),
                "u8": (capi.FORMAT_R8G8B8A8_UNORM, (np.clip(np.nan_to_num(a), 0, 1) * 255).astype(np.uint8))}[dtype]
    ctx = Context(0)
    img = DescriptorImage.create(ctx, fmt, W, H)
    img.compile()
    img.upload(np.ascontiguousarray(src))
    conv = FormatConverter.create(img, capi.FORMAT_B8G8R8A8_UNORM)
    conv.compile_images(ctx)
    conv.update_image_layouts(ctx)
    commands = Commands.create()
    conv.add_dispatch_to_command_graph(commands)
    commands.record()
    ctx.synchronize()
    np.testing.assert_array_equal(conv.final_image.download(), oracle.format_converter(np.ascontiguousarray(src)))
    with pytest.raises(VkpbrtError, match="Unknown format"):      # FormatConverter.cpp:13-19
        FormatConverter.create(img, capi.FORMAT_R8G8B8A8_UNORM)


def _demod_inputs(H, W, seed):
    rng = np.random.default_rng(seed)
    L = rng.uniform(-1, 14, (H, W, 4)).astype(np.float32)
    alb = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
    alb[2, :10, :3] = 0.0                                      # black albedo: the 1e3 cap
    px = rng.uniform(-50, 50, (H, W)).astype(np.float32)
    px[5:9, 20:40] = np.inf                                    # misses
    px[10, 3] = -np.inf
    sp = rng.random((H, W, 4)) < 0.05                          # zeros, denormals, huge values, infinities
    L[sp] = rng.choice(np.array([0.0, -0.0, 1e-40, 1e-6, 9.999999, 10.0, 10.000001, 3e38, np.inf, -np.inf], np.float32), size=int(sp.sum()))
    sp = rng.random((H, W, 4)) < 0.05
    alb[sp] = rng.choice(np.array([0.0, 1e-7, 1e-6, 1e-3, 0.01, 1.0], np.float32), size=int(sp.sum()))
    return L, alb, px


@pytest.mark.parametrize("H,W", [(33, 70), (1, 1), (16, 17)])
def test_demodulate_oracle_equals_reference_statements(oracle, H, W):
    """pins vkpbrt_oracle_demodulate: the reference's own statements (ptRaygen.rgen:81-88 -- the radiance clamp and the
    DEMOD_ILLUMINATION_FLOAT block, cut out of the shader's text and wrapped in a compute main() by
    oracle/glsl_shim/extract_rgen.py) against the oracle, bit for bit"""
    from oracle import ref as R
    if not R.build():
        pytest.skip("oracle/_ref is not built and /root/reference is not mounted")
    L, alb, px = _demod_inputs(max(H, 12), max(W, 41), 5)
    L, alb, px = np.ascontiguousarray(L[:H, :W]), np.ascontiguousarray(alb[:H, :W]), np.ascontiguousarray(px[:H, :W])
    np.testing.assert_array_equal(oracle.demodulate(L, alb, px).view(np.uint32), R.demodulate(L, alb, px).view(np.uint32))


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_demodulate(backend, oracle):
    """min(clamp(L, 0, 10) / (albedo + 1e-6), 1e3) for hits, the clamped radiance for misses (ptRaygen.rgen:81-88)"""
    H, W = 33, 70
    rng = np.random.default_rng(5)
    L = rng.uniform(-1, 14, (H, W, 4)).astype(np.float32)
    alb = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
    alb[2, :10, :3] = 0.0                                      # black albedo: the 1e3 cap
    px = rng.uniform(-50, 50, (H, W)).astype(np.float32)
    px[5:9, 20:40] = np.inf                                    # misses
    px[10, 3] = -np.inf
    ctx = Context(0)
    mk = lambda fmt, a: (lambda im: (im.compile(), im.upload(np.ascontiguousarray(a)), im)[2])(DescriptorImage.create(ctx, fmt, W, H))
    iL, ia, ip = mk(capi.FORMAT_R32G32B32A32_SFLOAT, L), mk(capi.FORMAT_R32G32B32A32_SFLOAT, alb), mk(capi.FORMAT_R32_SFLOAT, px)
    out = DescriptorImage.create(ctx, capi.FORMAT_R32G32B32A32_SFLOAT, W, H)
    out.compile()
    demodulate(ctx, iL, ia, ip, out)
    ctx.synchronize()
    want = oracle.demodulate(L, alb, px)
    np.testing.assert_array_equal(out.download().view(np.uint32), want.view(np.uint32))
    assert want[2, 0, 0] in (1e3, np.float32(1e3)) or L[2, 0, 0] <= 0
    np.testing.assert_array_equal(want[6, 25, :3], np.clip(L[6, 25, :3], 0, 10))


@pytest.mark.parametrize("backend", [backend_params()[0]], indirect=True)
def test_demodulate_special_values(backend, oracle):
    """zeros, denormals, values at the clamp, huge values and infinities in radiance / albedo (emulator only: added after
    the round's last GPU visit)"""
    H, W = 33, 70
    L, alb, px = _demod_inputs(H, W, 9)
    ctx = Context(0)
    mk = lambda fmt, a: (lambda im: (im.compile(), im.upload(np.ascontiguousarray(a)), im)[2])(DescriptorImage.create(ctx, fmt, W, H))
    iL, ia, ip = mk(capi.FORMAT_R32G32B32A32_SFLOAT, L), mk(capi.FORMAT_R32G32B32A32_SFLOAT, alb), mk(capi.FORMAT_R32_SFLOAT, px)
    out = DescriptorImage.create(ctx, capi.FORMAT_R32G32B32A32_SFLOAT, W, H)
    out.compile()
    demodulate(ctx, iL, ia, ip, out)
    ctx.synchronize()
    np.testing.assert_array_equal(out.download().view(np.uint32), oracle.demodulate(L, alb, px).view(np.uint32))
