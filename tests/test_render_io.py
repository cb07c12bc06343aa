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


def _cxx_offline(tmp_path, backend):
    """examples/cpp_offline_sequence.cpp built against the product library (cuda) or the emulator (hostsim)"""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    libdir, lib = (root / "tests" / "hostsim", "vkpbrt_hostsim") if backend == "hostsim" else (root / "vulkanpbrt_b200" / "lib", "vkpbrt_b200")
    exe = tmp_path / "cpp_offline_sequence"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", str(root / "include"), str(root / "examples" / "cpp_offline_sequence.cpp"), "-o", str(exe),
                        f"-L{libdir}", f"-l{lib}", f"-Wl,-rpath,{libdir}", "-lz"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("source", ["position", "depth"])
def test_cxx_offline_sequence_import(tmp_path, backend, source):
    """include/vkpbrt/io.hpp (MatrixIO, GBufferIO, IlluminationBufferIO, the EXR reader) in the reference's offline
    frame loop (examples/cpp_offline_sequence.cpp): a sequence exported by the Python layer (OpenCV's OpenEXR, ZIP, and
    the reference's matrix JSON) is imported by the C++ layer and denoised; the imported G-buffer and every final image
    equal the Python layer's import of the same files bit for bit (both run the conversions on the device)"""
    import subprocess
    from vulkanpbrt_b200 import DenoisePipeline, _capi as capi
    from vulkanpbrt_b200.matrix_io import export_matrices, import_matrices
    W, H, frames = 192, 128, 3
    seq = _sequence(W, H, frames)
    d = str(tmp_path)
    g = [OfflineGBuffer(depth=fr.depth, normal=fr.normal, material=fr.material, albedo=fr.albedo) for fr in seq]
    mats = [CameraMatrices(view=fr.camera.view, inv_view=fr.camera.inv_view, proj=fr.camera.proj, inv_proj=fr.camera.inv_proj) for fr in seq]
    assert GBufferIO.export_g_buffer(d + "/pos_%d.exr", d + "/depth_%d.exr", d + "/normal_%d.exr", "", d + "/albedo_%d.exr", frames, g, mats)
    assert IlluminationBufferIO.export_illumination(d + "/illu_%d.exr", frames, [OfflineIllumination(noisy=fr.illumination) for fr in seq])
    assert export_matrices(d + "/matrices.json", mats)
    exe = _cxx_offline(tmp_path, backend)
    out = tmp_path / "exported"
    out.mkdir()
    # the export flags of the reference's loop as well -- on the emulator only (memory copies + host loops, nothing of it computes on the device)
    r = subprocess.run([str(exe), d, str(frames), source] + ([str(out)] if backend == "hostsim" else []), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    if backend == "hostsim":
        # what the C++ layer exported is the imported G-buffer run through the oracle's export conversions, and imports again
        from oracle import oracle as orc        # the checker
        back = import_matrices(out / "matrices.json")
        for f in range(frames):
            dep = np.fromfile(tmp_path / f"gbuffer_{f}.depth", np.float32).reshape(H, W)
            nrm = np.fromfile(tmp_path / f"gbuffer_{f}.normal", np.float32).reshape(H, W, 2)
            alb = np.fromfile(tmp_path / f"gbuffer_{f}.albedo", np.uint8).reshape(H, W, 4)
            want_p, want_n, want_a = orc.gbuffer_export(np.asarray(mats[f].inv_view, np.float32), np.asarray(mats[f].inv_proj, np.float32), dep, nrm, alb)
            for name, want in (("pos", want_p), ("normal", want_n), ("albedo", want_a), ("depth", dep), ("illu", np.ascontiguousarray(seq[f].illumination, np.float32))):
                np.testing.assert_array_equal(read_exr(out / f"{name}_{f}.exr").reshape(want.shape).view(np.uint32), want.view(np.uint32), err_msg=f"exported {name} {f}")
            np.testing.assert_array_equal(np.asarray(back[f].view, np.float32), np.asarray(mats[f].view, np.float32))
    # the same import through the Python layer
    loaded = import_matrices(d + "/matrices.json")
    pipe = DenoisePipeline(W, H, use_taa=True, separate_matrices=True)
    for f in range(frames):
        pos = read_exr(d + f"/pos_{f}.exr") if source == "position" else None
        GBufferIO.import_to_device(pipe.g_buffer, pos, loaded[f] if pos is not None else None, read_exr(d + f"/normal_{f}.exr"), read_exr(d + f"/albedo_{f}.exr"))
        if source == "depth":
            pipe.g_buffer.depth.upload(read_exr(d + f"/depth_{f}.exr").reshape(H, W))
        pipe.raw_illumination.illumination_images[0].upload(np.ascontiguousarray(read_exr(d + f"/illu_{f}.exr")))
        np.testing.assert_array_equal(np.fromfile(tmp_path / f"gbuffer_{f}.depth", np.uint32).reshape(H, W), pipe.g_buffer.depth.download().view(np.uint32), err_msg=f"depth {f}")
        np.testing.assert_array_equal(np.fromfile(tmp_path / f"gbuffer_{f}.normal", np.uint32).reshape(H, W, 2), pipe.g_buffer.normal.download().view(np.uint32), err_msg=f"normal {f}")
        np.testing.assert_array_equal(np.fromfile(tmp_path / f"gbuffer_{f}.albedo", np.uint8).reshape(H, W, 4), pipe.g_buffer.albedo.download(), err_msg=f"albedo {f}")
        cur, prev = loaded[f], loaded[f - 1] if f > 0 else loaded[f]
        pc = pipe.push_constants.value
        pc.view_inverse = capi.mat16(cur.inv_view)
        pc.proj_inverse = capi.mat16(cur.inv_proj)
        pc.frame_number, pc.sample_number = f, 0
        pipe.accumulator.set_camera_matrices(f, cur, prev)
        pipe.record()
        pc.prev_view = capi.mat16(cur.view)
        pipe.ctx.synchronize()
        np.testing.assert_array_equal(np.fromfile(tmp_path / f"final_{f}.bgra", np.uint8).reshape(H, W, 4), pipe.final.download(), err_msg=f"final {f}")
    # with positions the planes are what the source frames held (up to the stored precision), not merely self-consistent
    assert np.abs(pipe.g_buffer.albedo.download().astype(int) - seq[-1].albedo.astype(int)).max() <= 1


def _exr_offset_table(b: bytes) -> int:
    """byte position of the scan line offset table: right after the header's empty attribute name"""
    i = 8                                             # magic + version
    while b[i] != 0:
        i = b.index(b"\0", i) + 1                     # attribute name
        i = b.index(b"\0", i) + 1                     # type name
        i += 4 + int.from_bytes(b[i:i + 4], "little")
    return i + 1


def test_cxx_exr_reader_and_matrix_files(tmp_path):
    """io.hpp on its own (no device): EXR files written by OpenCV in float / half, 1 / 3 / 4 channels, ZIP (default), and its
    own uncompressed files read back by OpenCV; the matrix JSON and the BMFR-dataset text layout against the Python reader"""
    import subprocess
    from pathlib import Path
    from vulkanpbrt_b200.matrix_io import export_matrices, import_matrices
    root = Path(__file__).resolve().parents[1]
    src = tmp_path / "io_probe.cpp"
    src.write_text(r'''
#include <vkpbrt/io.hpp>
using namespace vkpbrt;
int main(int argc, char** argv) {
    const std::string dir = argv[1];
    for (const char* name : {"rgba_f32", "rgb_f32", "gray_f32", "rgba_f16", "big_zip"}) {
        exr::Image im;
        if (!exr::read(dir + "/" + name + ".exr", im)) return 1;
        std::ofstream(dir + "/" + name + ".raw", std::ios::binary).write((const char*)im.data.data(), im.data.size() * 4);
        std::ofstream(dir + "/" + name + ".dims") << im.width << " " << im.height << " " << im.channels;
        if (!exr::write(dir + "/" + name + "_back.exr", im.data.data(), im.width, im.height, im.channels)) return 2;
    }
    exr::Image none;
    if (exr::read(dir + "/missing.exr", none) || exr::read(dir + "/garbage.exr", none)) return 3;
    if (exr::read(dir + "/huge_window.exr", none) || exr::read(dir + "/wild_offset.exr", none)) return 7;
    auto m = MatrixIO::import_matrices(dir + "/m.json");
    if (!MatrixIO::export_matrices(dir + "/m_back.json", m)) return 4;
    auto t = MatrixIO::import_matrices(dir + "/cams.txt");
    if (!MatrixIO::export_matrices(dir + "/t_back.json", t)) return 5;
    if (!MatrixIO::import_matrices(dir + "/nope.json").empty()) return 6;
    return 0;
}
''')
    exe = tmp_path / "io_probe"
    libdir = root / "tests" / "hostsim"
    subprocess.run(["make", "-C", str(libdir)], check=True, capture_output=True)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", str(root / "include"), str(src), "-o", str(exe), f"-L{libdir}", "-lvkpbrt_hostsim",
                        f"-Wl,-rpath,{libdir}", "-lz"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rng = np.random.default_rng(11)
    planes = {"rgba_f32": rng.standard_normal((9, 13, 4)).astype(np.float32), "rgb_f32": rng.standard_normal((7, 5, 3)).astype(np.float32),
              "gray_f32": rng.standard_normal((33, 17)).astype(np.float32), "big_zip": np.tile(rng.standard_normal((1, 70, 4)).astype(np.float32), (50, 1, 1))}
    for name, a in planes.items():
        assert write_exr(tmp_path / f"{name}.exr", a)
    half = rng.standard_normal((6, 10, 4)).astype(np.float16)
    b = np.ascontiguousarray(half.astype(np.float32)[..., [2, 1, 0, 3]])
    assert cv2.imwrite(str(tmp_path / "rgba_f16.exr"), b, [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_HALF])
    planes["rgba_f16"] = half.astype(np.float32)
    (tmp_path / "garbage.exr").write_bytes(b"not an exr file at all")
    # corrupt sizes must be refused, not allocated or dereferenced (found by mutating files under ASan): a data window of 2^31
    # columns, and a scan line offset near 2^64
    good = bytearray((tmp_path / "rgba_f32.exr").read_bytes())
    at = good.index(b"dataWindow\0box2i\0") + len(b"dataWindow\0box2i\0") + 4
    huge = bytearray(good)
    huge[at + 8:at + 12] = (0x7FFFFFFF).to_bytes(4, "little")
    (tmp_path / "huge_window.exr").write_bytes(huge)
    wild = bytearray(good)
    table = _exr_offset_table(good)
    wild[table:table + 8] = (0xFFFFFFFFFFFFFFFC).to_bytes(8, "little")
    (tmp_path / "wild_offset.exr").write_bytes(wild)
    seq = _sequence(32, 32, 2)
    mats = [CameraMatrices(view=fr.camera.view, inv_view=fr.camera.inv_view, proj=fr.camera.proj, inv_proj=fr.camera.inv_proj) for fr in seq]
    mats.append(CameraMatrices(view=seq[0].camera.view, inv_view=seq[0].camera.inv_view))           # one combined-matrix entry
    assert export_matrices(tmp_path / "m.json", mats)
    (tmp_path / "cams.txt").write_text("{" + ", ".join(f"{v:.9g}" for v in seq[0].camera.view) + "},\n{" + " ".join(f"{v:.9g}," for v in seq[1].camera.proj) + "}\n")
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "missing.exr" not in r.stderr and "garbage.exr" in r.stderr and "huge_window.exr" in r.stderr and "wild_offset.exr" in r.stderr
    for name, a in planes.items():
        w, h, c = map(int, (tmp_path / f"{name}.dims").read_text().split())
        assert (h, w) == a.shape[:2] and c == (a.shape[2] if a.ndim == 3 else 1)
        got = np.fromfile(tmp_path / f"{name}.raw", np.float32).reshape(a.shape)
        np.testing.assert_array_equal(got.view(np.uint32), a.view(np.uint32), err_msg=name)
        back = read_exr(tmp_path / f"{name}_back.exr")                                              # io.hpp's writer -> OpenCV's reader
        np.testing.assert_array_equal(back.reshape(a.shape).view(np.uint32), a.view(np.uint32), err_msg=name + " (written by io.hpp)")
    for back, want in (("m_back.json", mats), ("t_back.json", import_matrices(tmp_path / "cams.txt"))):
        got = import_matrices(tmp_path / back)
        assert len(got) == len(want) and len(want) >= 2
        for x, y in zip(got, want):
            np.testing.assert_array_equal(np.asarray(x.view, np.float32), np.asarray(y.view, np.float32))
            np.testing.assert_array_equal(np.asarray(x.inv_view, np.float32), np.asarray(y.inv_view, np.float32))
            assert (x.proj is None) == (y.proj is None)
            if y.proj is not None:
                np.testing.assert_array_equal(np.asarray(x.inv_proj, np.float32), np.asarray(y.inv_proj, np.float32))


def test_import_oracle_equals_the_reference_host_code(oracle):
    """pins vkpbrt_oracle_gbuffer_import: the reference's own C++ (GBufferIO::convert_normal_to_spherical,
    GBufferIO::compress_albedo and the position -> depth block of import_g_buffer_position, cut out of
    source/io/RenderIO.cpp by oracle/host_shim/extract_host.py and compiled against vsg's maths headers) against the
    oracle: every plane bit for bit -- this is what showed that `camera_pos /= camera_pos.w` multiplies by the
    reciprocal in vsg, and that the unqualified acos / atan2 of :174-175 are the C library's double routines (only
    <cmath> is in scope), rounded once by the store"""
    from oracle import ref as R
    if not R.build_host():
        pytest.skip("oracle/_ref host library is not built and /root/reference is not mounted")
    rng = np.random.default_rng(3)
    for (H, W) in ((37, 50), (1, 1), (16, 129)):
        pos = rng.uniform(-50, 50, (H, W, 4)).astype(np.float32)
        v = rng.standard_normal((H, W, 3))
        v /= np.linalg.norm(v, axis=-1, keepdims=True)
        nrm = np.concatenate([v, np.ones((H, W, 1))], -1).astype(np.float32)
        special = np.array([[0, 0, 1, 1], [0, 0, -1, 1], [1, 0, 0, 1], [0, -1, 0, 1], [-1, 0, 0, 1], [0.6, 0.8, 0, 1]], np.float32)
        nrm.reshape(-1, 4)[:min(len(special), H * W)] = special[:H * W]
        alb = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
        edge = np.array([0, 0.5, 1.0, 0.999, 0.0039, 0.25, 1 / 255, 254.5 / 255], np.float32)
        alb.reshape(-1)[:min(len(edge), alb.size)] = edge[:alb.size]
        for trial in range(4):
            iv = rng.uniform(-2, 2, 16).astype(np.float32)
            iv[11] = np.float32([0.37, 3.0, -1.7, 1.0][trial])           # the w the eye point is divided by
            a, b = oracle.gbuffer_import(iv, pos, nrm, alb), R.gbuffer_import(iv, pos, nrm, alb)
            np.testing.assert_array_equal(a[0].view(np.uint32), b[0].view(np.uint32), err_msg="depth")
            np.testing.assert_array_equal(a[2], b[2], err_msg="albedo")
            np.testing.assert_array_equal(a[1].view(np.uint32), b[1].view(np.uint32), err_msg="normal")


def _export_inputs(rng, H, W):
    depth = rng.uniform(0.05, 200, (H, W)).astype(np.float32)
    depth.reshape(-1)[:3] = np.float32([0.0, 1e10, np.inf])[:min(3, H * W)]             # a hit at the eye, the miss distance, an overflowed one
    sph = np.stack([rng.uniform(0, np.pi, (H, W)), rng.uniform(-np.pi, np.pi, (H, W))], -1).astype(np.float32)
    special = np.float32([[0, 0], [np.pi, 0], [np.pi / 2, np.pi], [0, np.pi / 4], [1e-30, -np.pi / 2], [3.0, 100.0]])    # poles, the miss normal, a denormal-ish angle, phi beyond pi
    sph.reshape(-1, 2)[:min(len(special), H * W)] = special[:H * W]
    unorm = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    unorm.reshape(-1)[:min(4, unorm.size)] = np.uint8([0, 1, 254, 255])[:unorm.size]
    return depth, sph, unorm


def test_export_oracle_equals_the_reference_host_code(oracle):
    """pins vkpbrt_oracle_gbuffer_export: GBufferIO::depth_to_position, spherical_to_cartesian and unorm_to_float, cut whole out
    of source/io/RenderIO.cpp (:312-382) by oracle/host_shim/extract_host.py and compiled against vsg's maths headers, against
    the oracle -- every plane bit for bit (this is what showed that the unqualified cos / sin of :322-324 are the double
    routines, and that normalize() multiplies by the reciprocal of the length), and no position plane without a separate
    projection matrix"""
    from oracle import ref as R
    if not R.build_host():
        pytest.skip("oracle/_ref host library is not built and /root/reference is not mounted")
    rng = np.random.default_rng(17)
    for (H, W) in ((37, 50), (1, 1), (16, 129), (3, 2)):
        depth, sph, unorm = _export_inputs(rng, H, W)
        for trial in range(3):
            m = rng.uniform(-2, 2, 64).astype(np.float32)
            if trial == 2:                                                                # a real camera
                c = synth.render_frame(32, 32, 5).camera
                m = np.concatenate([np.asarray(x, np.float32).reshape(-1) for x in (c.view, c.inv_view, c.proj, c.inv_proj)])
            a = oracle.gbuffer_export(m[16:32], m[48:64], depth, sph, unorm)
            b = R.gbuffer_export(m, True, depth, sph, unorm)
            for got, want, name in zip(a, b, ("position", "normal", "unorm")):
                np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32), err_msg=f"{name} {H}x{W} trial {trial}")
    assert R.gbuffer_export(m, False, depth)[0] is None and oracle.gbuffer_export(m[16:32], None, depth)[0] is None


def test_export_conversions_equal_the_oracle(oracle):
    """the Python layer's export conversions against the oracle: positions and unorm -> float bit for bit, cartesian normals
    within one ulp (numpy's double cos / sin are not the C library's)"""
    rng = np.random.default_rng(23)
    H, W = 41, 67
    depth, sph, unorm = _export_inputs(rng, H, W)
    c = synth.render_frame(W, H, 3).camera
    cm = CameraMatrices(view=c.view, inv_view=c.inv_view, proj=c.proj, inv_proj=c.inv_proj)
    want_p, want_n, want_u = oracle.gbuffer_export(np.asarray(c.inv_view, np.float32), np.asarray(c.inv_proj, np.float32), depth, sph, unorm)
    np.testing.assert_array_equal(GBufferIO.depth_to_position(depth, cm).view(np.uint32), want_p.view(np.uint32))
    np.testing.assert_array_equal(GBufferIO.unorm_to_float(unorm).view(np.uint32), want_u.view(np.uint32))
    np.testing.assert_array_max_ulp(GBufferIO.spherical_to_cartesian(sph), want_n, maxulp=1)


def test_cxx_g_buffer_export(tmp_path, oracle):
    """include/vkpbrt/io.hpp's export side on the emulator (memory copies only: nothing here computes on the device): planes
    uploaded into a GBuffer / an illumination buffer are read back by OfflineGBuffer::download_from_g_buffer and
    OfflineIllumination::download_from_illumination_buffer, written by GBufferIO::export_g_buffer /
    IlluminationBufferIO::export_illumination, and the files -- decoded by OpenCV -- hold the oracle's conversions bit for
    bit; a skipped plane (empty format) writes nothing, an unwritable path and a missing projection matrix fail the call"""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    src = tmp_path / "export_probe.cpp"
    src.write_text(r'''
#include <vkpbrt/io.hpp>
using namespace vkpbrt;
template <class T> static std::vector<T> slurp(const std::string& path)
{
    std::ifstream f(path, std::ios::binary);
    std::vector<char> b((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<T> out(b.size() / sizeof(T));
    std::memcpy(out.data(), b.data(), out.size() * sizeof(T));
    return out;
}
int main(int argc, char** argv) {
    const std::string dir = argv[1];
    const uint32_t W = std::stoul(argv[2]), H = std::stoul(argv[3]);
    const int frames = std::stoi(argv[4]);
    auto ctx = Context::create(0);
    auto matrices = MatrixIO::import_matrices(dir + "/matrices.json");
    auto g_buffer = GBuffer::create(*ctx, W, H);
    g_buffer->compile(*ctx);
    ref_ptr<IlluminationBuffer> illu = IlluminationBufferDemodulatedFloat::create(*ctx, W, H);
    illu->compile(*ctx);
    OfflineGBuffers g(frames);
    OfflineIlluminations il(frames);
    for (int f = 0; f < frames; ++f) {
        const std::string n = std::to_string(f);
        auto d = slurp<float>(dir + "/in_depth_" + n), nr = slurp<float>(dir + "/in_normal_" + n), no = slurp<float>(dir + "/in_illu_" + n);
        auto al = slurp<uint8_t>(dir + "/in_albedo_" + n), ma = slurp<uint8_t>(dir + "/in_material_" + n);
        check(vkpbrt_image_upload(g_buffer->depth->handle, d.data(), d.size() * 4));
        check(vkpbrt_image_upload(g_buffer->normal->handle, nr.data(), nr.size() * 4));
        check(vkpbrt_image_upload(g_buffer->albedo->handle, al.data(), al.size()));
        check(vkpbrt_image_upload(g_buffer->material->handle, ma.data(), ma.size()));
        check(vkpbrt_image_upload(illu->illumination_images[0]->handle, no.data(), no.size() * 4));
        g[f] = OfflineGBuffer::create();
        g[f]->download_from_g_buffer(g_buffer, *ctx);
        il[f] = OfflineIllumination::create();
        il[f]->download_from_illumination_buffer(illu, *ctx);
    }
    if (!GBufferIO::export_g_buffer(dir + "/pos_%d.exr", dir + "/depth_%d.exr", dir + "/normal_%d.exr", dir + "/material_%d.exr", dir + "/albedo_%d.exr", frames, g, matrices, 0)) return 1;
    if (!IlluminationBufferIO::export_illumination(dir + "/illu_%d.exr", frames, il, 0)) return 2;
    if (!GBufferIO::export_g_buffer("", dir + "/only_depth_%d.exr", "", "", "", frames, g, matrices, 0)) return 3;
    if (GBufferIO::export_g_buffer("", dir + "/no/such/dir/depth_%d.exr", "", "", "", frames, g, matrices, 0)) return 4;
    std::vector<CameraMatrices> combined(matrices);
    for (auto& m : combined) { m.proj.reset(); m.inv_proj.reset(); }
    if (GBufferIO::export_g_buffer(dir + "/nopos_%d.exr", "", "", "", "", frames, g, combined, 0)) return 5;
    return 0;
}
''')
    from vulkanpbrt_b200.matrix_io import export_matrices
    libdir = root / "tests" / "hostsim"
    subprocess.run(["make", "-C", str(libdir)], check=True, capture_output=True)
    exe = tmp_path / "export_probe"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-I", str(root / "include"), str(src), "-o", str(exe), f"-L{libdir}", "-lvkpbrt_hostsim",
                        f"-Wl,-rpath,{libdir}", "-lz"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    W, H, frames = 80, 48, 2
    seq = _sequence(W, H, frames)
    mats = [CameraMatrices(view=fr.camera.view, inv_view=fr.camera.inv_view, proj=fr.camera.proj, inv_proj=fr.camera.inv_proj) for fr in seq]
    assert export_matrices(tmp_path / "matrices.json", mats)
    rng = np.random.default_rng(4)
    material = [rng.integers(0, 256, (H, W, 4), dtype=np.uint8) for _ in seq]
    for f, fr in enumerate(seq):
        for name, a in (("depth", fr.depth), ("normal", fr.normal), ("illu", fr.illumination), ("albedo", fr.albedo), ("material", material[f])):
            np.ascontiguousarray(a).tofile(tmp_path / f"in_{name}_{f}")
    r = subprocess.run([str(exe), str(tmp_path), str(W), str(H), str(frames)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "no/such/dir" in r.stderr and "depthToPosition" in r.stdout
    assert not list(tmp_path.glob("nopos_*")) and len(list(tmp_path.glob("only_depth_*.exr"))) == frames
    for f, fr in enumerate(seq):
        c = fr.camera
        want_p, want_n, want_a = oracle.gbuffer_export(np.asarray(c.inv_view, np.float32), np.asarray(c.inv_proj, np.float32), fr.depth, fr.normal, fr.albedo)
        want_m = oracle.gbuffer_export(unorm=material[f])[2]
        for name, want in (("pos", want_p), ("normal", want_n), ("albedo", want_a), ("material", want_m), ("illu", np.ascontiguousarray(fr.illumination, np.float32)),
                           ("depth", fr.depth), ("only_depth", fr.depth)):
            got = read_exr(tmp_path / f"{name}_{f}.exr")
            np.testing.assert_array_equal(got.reshape(want.shape).view(np.uint32), want.view(np.uint32), err_msg=f"{name} {f}")
        # and the Python layer's export of the same frame holds the same planes
    g = [OfflineGBuffer(depth=fr.depth, normal=fr.normal, material=material[f], albedo=fr.albedo) for f, fr in enumerate(seq)]
    d = str(tmp_path)
    assert GBufferIO.export_g_buffer(d + "/py_pos_%d.exr", "", "", d + "/py_material_%d.exr", d + "/py_albedo_%d.exr", frames, g, mats, verbosity=0)
    for f in range(frames):
        for name in ("pos", "material", "albedo"):
            np.testing.assert_array_equal(read_exr(tmp_path / f"py_{name}_{f}.exr").view(np.uint32), read_exr(tmp_path / f"{name}_{f}.exr").view(np.uint32), err_msg=f"py {name} {f}")


def test_seeded_fuzz_of_the_host_conversions_against_the_reference(oracle):
    """oracle against the reference's host code on extreme data: magnitudes over 18 decades, zeros, denormals, infinities and
    NaNs sprinkled over positions / normals / depths / angles / matrices (300 draws of this ran clean by hand); every plane
    bit for bit with NaNs compared as NaNs.  Albedo stays in [0, 1]: the reference's float -> byte conversion of anything else
    is undefined behaviour"""
    from oracle import ref as R
    if not R.build_host():
        pytest.skip("oracle/_ref host library is not built and /root/reference is not mounted")
    rng = np.random.default_rng(99)
    spec = np.float32([0, -0.0, 1e-45, 1e-38, 1e-20, 1e20, 3e38, np.inf, -np.inf, np.nan, -1, 1, np.pi, -np.pi, np.pi / 2])

    def same(a, b, what):
        a, b = a.view(np.uint32).copy(), b.view(np.uint32).copy()
        na, nb = (a & 0x7FFFFFFF) > 0x7F800000, (b & 0x7FFFFFFF) > 0x7F800000
        np.testing.assert_array_equal(na, nb, err_msg=what + " (NaN positions)")
        a[na] = 0
        b[nb] = 0
        np.testing.assert_array_equal(a, b, err_msg=what)

    def sprinkle(arr, p):
        k = rng.random(arr.shape) < p
        arr[k] = rng.choice(spec, size=int(k.sum()))

    for it in range(40):
        H, W = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        depth = (10.0 ** rng.uniform(-6, 12, (H, W))).astype(np.float32)
        sph = rng.uniform(-10, 10, (H, W, 2)).astype(np.float32)
        m = (rng.uniform(-2, 2, 64) * 10.0 ** rng.uniform(-3, 3)).astype(np.float32)
        pos = (rng.uniform(-50, 50, (H, W, 4)) * 10.0 ** rng.uniform(-3, 6)).astype(np.float32)
        nrm = rng.uniform(-1.2, 1.2, (H, W, 4)).astype(np.float32)                    # |n.z| > 1: acos -> NaN on both sides
        alb = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
        for arr, p in ((depth, 0.08), (sph, 0.08), (m, 0.08), (pos, 0.05), (nrm, 0.05)):
            sprinkle(arr, p)
        unorm = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        a, b = oracle.gbuffer_export(m[16:32], m[48:64], depth, sph, unorm), R.gbuffer_export(m, True, depth, sph, unorm)
        for x, y, name in zip(a, b, ("position", "cartesian normal", "unorm")):
            same(x, y, f"export {name}, draw {it}")
        c, d = oracle.gbuffer_import(m[16:32], pos, nrm, alb), R.gbuffer_import(m[16:32], pos, nrm, alb)
        same(c[0], d[0], f"import depth, draw {it}")
        same(c[1], d[1], f"import normal, draw {it}")
        np.testing.assert_array_equal(c[2], d[2], err_msg=f"import albedo, draw {it}")


@pytest.mark.parametrize("backend", [backend_params()[0]], indirect=True)
def test_download_helpers_feed_the_export(tmp_path, backend, oracle):
    """OfflineGBuffer.download_from_g_buffer / OfflineIllumination.download_from_illumination_buffer (the reference's export
    staging, RenderIO.cpp:401-467, :748-928) -> export_g_buffer / export_illumination -> the files hold the oracle's
    conversions of what was uploaded.  Emulator only: uploads, downloads and host code"""
    from vulkanpbrt_b200 import DenoisePipeline
    W, H = 64, 48
    fr = _sequence(W, H, 2)[1]
    pipe = DenoisePipeline(W, H, use_taa=False)
    pipe.upload_frame(fr)
    g = OfflineGBuffer().download_from_g_buffer(pipe.g_buffer)
    il = OfflineIllumination().download_from_illumination_buffer(pipe.raw_illumination)
    np.testing.assert_array_equal(g.depth, fr.depth)
    np.testing.assert_array_equal(g.normal, fr.normal)
    np.testing.assert_array_equal(g.albedo, fr.albedo)
    np.testing.assert_array_equal(il.noisy.view(np.uint32), np.ascontiguousarray(fr.illumination, np.float32).view(np.uint32))
    c = fr.camera
    cm = CameraMatrices(view=c.view, inv_view=c.inv_view, proj=c.proj, inv_proj=c.inv_proj)
    d = str(tmp_path)
    assert GBufferIO.export_g_buffer(d + "/pos_%d.exr", d + "/depth_%d.exr", d + "/normal_%d.exr", "", d + "/albedo_%d.exr", 1, [g], [cm], verbosity=0)
    assert IlluminationBufferIO.export_illumination(d + "/illu_%d.exr", 1, [il], verbosity=0)
    want_p, want_n, want_a = oracle.gbuffer_export(np.asarray(c.inv_view, np.float32), np.asarray(c.inv_proj, np.float32), fr.depth, fr.normal, fr.albedo)
    np.testing.assert_array_equal(read_exr(d + "/pos_0.exr").view(np.uint32), want_p.view(np.uint32))
    np.testing.assert_array_max_ulp(read_exr(d + "/normal_0.exr"), want_n, maxulp=1)
    np.testing.assert_array_equal(read_exr(d + "/albedo_0.exr").view(np.uint32), want_a.view(np.uint32))
    np.testing.assert_array_equal(read_exr(d + "/illu_0.exr").view(np.uint32), il.noisy.view(np.uint32))
