"""Offline sequence I/O (vulkanpbrt_b200/render_io.py ~ GBufferIO / IlluminationBufferIO, source/io/RenderIO.cpp)."""
import numpy as np
import pytest

from tests.conftest import backend_params
from vulkanpbrt_b200 import synth
from vulkanpbrt_b200.modules import CameraMatrices
from vulkanpbrt_b200.render_io import GBufferIO, IlluminationBufferIO, OfflineGBuffer, OfflineIllumination, read_exr, write_exr

cv2 = pytest.importorskip("cv2")


def test_exr_planes_keep_channel_order_and_bits(tmp_path):
    rng = np.random.default_rng(3)
    rgba = rng.standard_normal((9, 13, 4)).astype(np.float32)
    rgba[..., 0] += 10      # make a channel swap impossible to miss
    assert write_exr(tmp_path / "p.exr", rgba)
    back = read_exr(tmp_path / "p.exr")
    np.testing.assert_array_equal(back.view(np.uint32), rgba.view(np.uint32))
    gray = rng.standard_normal((9, 13)).astype(np.float32)
    assert write_exr(tmp_path / "d.exr", gray)
    np.testing.assert_array_equal(read_exr(tmp_path / "d.exr").reshape(9, 13), gray)
    assert read_exr(tmp_path / "missing.exr") is None


def test_normal_conversions_against_double_precision():
    rng = np.random.default_rng(5)
    v = rng.standard_normal((32, 48, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    n4 = np.concatenate([v, np.ones((32, 48, 1))], axis=-1).astype(np.float32)
    sph = GBufferIO.convert_normal_to_spherical(n4)
    assert sph.dtype == np.float32 and sph.shape == (32, 48, 2)
    np.testing.assert_allclose(sph[..., 0], np.arccos(n4[..., 2].astype(np.float64)), atol=2e-6)
    np.testing.assert_allclose(sph[..., 1], np.arctan2(n4[..., 1].astype(np.float64), n4[..., 0].astype(np.float64)), atol=2e-6)
    cart = GBufferIO.spherical_to_cartesian(sph)
    np.testing.assert_allclose(cart[..., :3], n4[..., :3], atol=3e-6)
    assert (cart[..., 3] == 1).all()


def test_albedo_compression_truncates_like_the_cxx_conversion():
    a = np.array([[[0.0, 0.5, 1.0, 1.0], [0.999, 0.0039, 0.25, 0.0]]], np.float32)
    got = GBufferIO.compress_albedo(a)
    assert got.dtype == np.uint8
    np.testing.assert_array_equal(got, [[[0, 127, 255, 255], [254, 0, 63, 0]]])          # 0.5 * 255 = 127.5 -> 127, not 128
    np.testing.assert_array_equal(GBufferIO.compress_albedo(got), got)                    # integer input: as is
    np.testing.assert_array_equal(GBufferIO.unorm_to_float(got), got.astype(np.float32) / np.float32(255))


def _sequence(W, H, frames):
    return [synth.render_frame(W, H, f) for f in range(frames)]


def test_position_round_trip_recovers_depth(tmp_path):
    """depth -> world position (separate matrices, export side) -> depth (import side)"""
    W, H = 96, 64
    fr = _sequence(W, H, 1)[0]
    c = fr.camera
    pos = GBufferIO.depth_to_position(fr.depth, CameraMatrices(view=c.view, inv_view=c.inv_view, proj=c.proj, inv_proj=c.inv_proj))
    assert pos.shape == (H, W, 4)
    # the import side reads the eye point from column 2 of the inverse of a COMBINED view-projection matrix
    # (RenderIO.cpp:109-110, the same convention as accumulator.comp:56-57)
    v = np.asarray(c.view, np.float64).reshape(4, 4).T
    p = np.asarray(c.proj, np.float64).reshape(4, 4).T
    ivp = np.linalg.inv(p @ v).T.astype(np.float32).reshape(-1)
    depth = GBufferIO.position_to_depth(pos, CameraMatrices(view=None, inv_view=ivp))
    hit = fr.depth < 1e4
    np.testing.assert_allclose(depth[hit], fr.depth[hit], rtol=2e-4)
    assert GBufferIO.depth_to_position(fr.depth, CameraMatrices(view=c.view, inv_view=c.inv_view)) is None


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_exported_sequence_imports_and_drives_the_modules(tmp_path, backend, oracle, capsys):
    """export a synthetic sequence the way the reference stores it (EXR: depth, cartesian normals, float albedo, 1-spp
    illumination), import it back and denoise it: every plane equals the oracle fed with the same imported planes"""
    from tests.util import assert_frame_equal, make_pair
    W, H, frames = 160, 96, 3
    seq = _sequence(W, H, frames)
    g = [OfflineGBuffer(depth=fr.depth, normal=fr.normal, material=fr.material, albedo=fr.albedo) for fr in seq]
    mats = [CameraMatrices(view=fr.camera.view, inv_view=fr.camera.inv_view, proj=fr.camera.proj, inv_proj=fr.camera.inv_proj) for fr in seq]
    d = str(tmp_path)
    assert GBufferIO.export_g_buffer(d + "/pos_%d.exr", d + "/depth_%d.exr", d + "/normal_%d.exr", "", d + "/albedo_%d.exr", frames, g, mats)
    assert IlluminationBufferIO.export_illumination(d + "/illu_%d.exr", frames, [OfflineIllumination(noisy=fr.illumination) for fr in seq])
    got = GBufferIO.import_g_buffer_depth(d + "/depth_%d.exr", d + "/normal_%d.exr", "", d + "/albedo_%d.exr", frames)
    illu = IlluminationBufferIO.import_illumination(d + "/illu_%d.exr", frames)
    assert len(got) == frames and len(illu) == frames
    for fr, gi, ii in zip(seq, got, illu):
        np.testing.assert_array_equal(gi.depth, fr.depth)                                    # float planes are stored losslessly
        np.testing.assert_array_equal(ii.noisy.view(np.uint32), fr.illumination.view(np.uint32))
        # through cartesian and back: phi is lost where theta = 0 (sky pixels), so compare directions
        np.testing.assert_allclose(GBufferIO.spherical_to_cartesian(gi.normal), GBufferIO.spherical_to_cartesian(fr.normal), atol=2e-6)
        assert np.abs(gi.albedo.astype(int) - fr.albedo.astype(int)).max() <= 1              # x/255*255 truncated
    pipe, orc = make_pair(oracle, W, H, denoiser="bmfr", block=32, use_taa=True)
    for f, fr in enumerate(seq):
        loaded = synth.Frame(f, got[f].depth, got[f].normal, got[f].albedo, got[f].material, illu[f].noisy, fr.camera)
        pipe.run_frame(f, loaded)
        pipe.ctx.synchronize()
        orc.run_frame(f, loaded)
        assert_frame_equal(pipe, orc, f)
    # a missing frame is reported like the reference does, and leaves that frame empty
    short = GBufferIO.import_g_buffer_depth(d + "/depth_%d.exr", d + "/normal_%d.exr", "", d + "/albedo_%d.exr", frames + 1)
    assert short[frames].depth is None and "Failed to load image" in capsys.readouterr().out


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_import_conversions_on_the_device(backend, oracle):
    """vkpbrt_gbuffer_import_record (k_gbuffer_import): GBufferIO's host conversions (RenderIO.cpp:101-120, :160-195) as one
    launch into a compiled GBuffer, against the oracle's restatement: depth and albedo bit for bit; the spherical normals
    within 1e-6 rad (acos / atan2 are the platform's libm in the reference, CUDA's on the device: a couple of ulps apart)"""
    from vulkanpbrt_b200 import Context
    from vulkanpbrt_b200.modules import GBuffer
    W, H = 70, 37
    fr = synth.render_frame(W, H, 2)
    c = fr.camera
    pos = GBufferIO.depth_to_position(fr.depth, CameraMatrices(view=c.view, inv_view=c.inv_view, proj=c.proj, inv_proj=c.inv_proj))
    v = np.asarray(c.view, np.float64).reshape(4, 4).T
    p = np.asarray(c.proj, np.float64).reshape(4, 4).T
    ivp = np.linalg.inv(p @ v).T.astype(np.float32).reshape(-1)          # combined inverse: the offline convention
    cart = GBufferIO.spherical_to_cartesian(fr.normal)
    alb = (fr.albedo.astype(np.float32) / np.float32(255.0) * np.float32(1.003)).astype(np.float32)      # some values above 1: the clamp
    ctx = Context(0)
    g = GBuffer.create(ctx, W, H)
    g.compile(ctx)
    GBufferIO.import_to_device(g, pos, CameraMatrices(view=None, inv_view=ivp), cart, alb)
    want_d, want_n, want_a = oracle.gbuffer_import(ivp, pos, cart, alb)
    np.testing.assert_array_equal(g.depth.download().view(np.uint32), want_d.view(np.uint32))
    np.testing.assert_array_equal(g.albedo.download(), want_a)
    np.testing.assert_allclose(g.normal.download(), want_n, atol=1e-6, rtol=0)
    # and the oracle's restatement is the host path the reference runs (numpy float32 == libm to the same tolerance)
    np.testing.assert_array_equal(want_d, GBufferIO.position_to_depth(pos, CameraMatrices(view=None, inv_view=ivp)))
    np.testing.assert_allclose(want_n, GBufferIO.convert_normal_to_spherical(cart), atol=1e-6, rtol=0)
    np.testing.assert_array_equal(want_a, GBufferIO.compress_albedo(alb))
