"""Camera-matrix files of the offline mode (vulkanpbrt_b200/matrix_io.py ~ MatrixIO, source/io/RenderIO.cpp:593-711)."""
import ctypes as C
import json

import numpy as np
import pytest

from tests.conftest import backend_params
from vulkanpbrt_b200 import synth
from vulkanpbrt_b200.matrix_io import _inverse, export_matrices, import_matrices
from vulkanpbrt_b200.modules import CameraMatrices


def _cams(n, W=320, H=200):
    return [synth.camera(W, H, f) for f in range(n)]


def test_json_round_trip_separate_matrices(tmp_path):
    cams = _cams(5)
    mats = [CameraMatrices(view=c.view, inv_view=c.inv_view, proj=c.proj, inv_proj=c.inv_proj) for c in cams]
    path = tmp_path / "camera.json"
    assert export_matrices(path, mats)
    doc = json.loads(path.read_text())
    # the reference's own keys (RenderIO.cpp:693-707)
    assert doc["amtOfFrames"] == 5 and set(doc["matrices"][0]) == {"type", "storageType", "view", "invView", "proj", "invProj"}
    assert doc["matrices"][0]["type"] == "ModelView+Projection" and doc["matrices"][0]["storageType"] == "ColumnMajor"
    back = import_matrices(path)
    assert len(back) == 5
    for c, m in zip(cams, back):
        for name in ("view", "inv_view", "proj", "inv_proj"):
            np.testing.assert_array_equal(getattr(m, name), getattr(c, name).astype(np.float32))


def test_json_combined_matrices_have_no_projection(tmp_path):
    c = _cams(1)[0]
    path = tmp_path / "vp.json"
    export_matrices(path, [CameraMatrices(view=c.view, inv_view=c.inv_view)])
    assert json.loads(path.read_text())["matrices"][0]["type"] == "ModelViewProjection"
    m = import_matrices(path)[0]
    assert m.proj is None and m.inv_proj is None
    assert m.to_c().has_proj == 0


def test_bmfr_dataset_text_format(tmp_path, oracle):
    """16 numbers per matrix, separated by whitespace, with the dataset's braces and commas (RenderIO.cpp:639-664)"""
    cams = _cams(3)
    lines = []
    for c in cams:
        v = [repr(float(x)) for x in c.view]
        lines.append("{" + ", ".join(v[:8]) + ",\n " + ", ".join(v[8:]) + "},")
    path = tmp_path / "camera_matrices.h"
    path.write_text("const float camera_matrices[3][4][4] = {\n" + "\n".join(lines) + "\n};\n")
    got = import_matrices(path)
    assert len(got) == 3
    L = oracle.lib()
    for c, m in zip(cams, got):
        np.testing.assert_array_equal(m.view, c.view.astype(np.float32))
        want = (C.c_float * 16)()
        L.vkpbrt_oracle_vsg_inverse((C.c_float * 16)(*[float(x) for x in c.view]), want)      # RenderIO.cpp:659 inverse(tmp) is vsg's
        np.testing.assert_array_equal(m.inv_view.view(np.uint32), np.array(list(want), np.float32).view(np.uint32))
        assert m.proj is None


def _matrices(n, seed):
    """affine, general and singular matrices, plus the synthetic sequence's own"""
    rng = np.random.default_rng(seed)
    out = []
    for f in range(12):
        cam = synth.camera(160, 128, f)
        v = np.asarray(cam.view, np.float64).reshape(4, 4).T
        p = np.asarray(cam.proj, np.float64).reshape(4, 4).T
        out += [cam.view, cam.proj, cam.inv_view, cam.inv_proj, (p @ v).T.astype(np.float32).reshape(-1)]
    for _ in range(n):
        m = rng.uniform(-3, 3, 16).astype(np.float32)
        if rng.random() < 0.5:
            m[3] = m[7] = m[11] = 0
            m[15] = 1                      # affine: vsg takes t_inverse_4x3
        if rng.random() < 0.05:
            m[4:8] = m[0:4]                # singular: NaN on the diagonal
        out.append(m)
    out += [np.zeros(16, np.float32), np.eye(4, dtype=np.float32).reshape(-1)]
    return [np.ascontiguousarray(m, np.float32) for m in out]


def _same_bits(a, b):
    x, y = np.asarray(a, np.float32).view(np.uint32).copy(), np.asarray(b, np.float32).view(np.uint32).copy()
    nan = np.isnan(a) & np.isnan(b)
    x[nan] = 0
    y[nan] = 0
    return np.array_equal(x, y)


def test_inverse_matches_the_oracle_bit_for_bit(oracle):
    """the library's vkpbrt_mat4_inverse (what matrix_io and include/vkpbrt/io.hpp call) against the oracle's restatement"""
    for m in _matrices(300, 7):
        assert _same_bits(_inverse(m), oracle.vsg_inverse(m))


def _hostref():
    from oracle import ref as R
    if not R.build_host():
        pytest.skip("oracle/_ref host library is not built and /root/reference is not mounted")
    h = C.CDLL(str(R._HOST_LIB))
    return h


def test_vsg_inverse_equals_the_reference_source(oracle):
    """pins vkpbrt_oracle_vsg_inverse: vsg's own t_inverse_4x3 / t_inverse_4x4 / inverse(mat4) text
    (external/vsg/src/vsg/maths/maths_transform.cpp, compiled by oracle/host_shim) on affine, general and singular
    matrices, bit for bit.  (Round 1 and most of round 2 restated the HOST inverse with the shader-side cofactor formula:
    same value to the last bits or so, not the same bits.)"""
    h = _hostref()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for m in _matrices(2000, 1):
        want = np.zeros(16, np.float32)
        h.hostref_inverse(p(m), p(want))
        assert _same_bits(oracle.vsg_inverse(m), want), m


@pytest.mark.parametrize("separate", [True, False])
def test_set_camera_matrices_equals_the_reference_source(oracle, separate):
    """pins the oracle's Accumulator::set_camera_matrices against the reference's own text (Accumulator.cpp:85-117, compiled
    by oracle/host_shim together with vsg's inverse): the 212-byte push-constant block frame by frame, both matrix modes;
    `inverse(prev.view)[3]` and `prev_pos /= prev_pos.w` (a multiplication by the reciprocal in vsg) included"""
    from vulkanpbrt_b200.pipeline import _combined
    h = _hostref()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    W, H = 160, 128
    chain = oracle.OracleChain(W, H, "bmfr", 32, separate_matrices=separate)
    pc = np.zeros(53, np.float32)

    def pack(cam):
        if separate:
            return np.concatenate([cam.view, cam.inv_view, cam.proj, cam.inv_proj]).astype(np.float32), 1
        vp, ivp = _combined(cam)
        return np.concatenate([vp, ivp, np.zeros(32, np.float32)]).astype(np.float32), 0

    prev_cam = None
    for f in range(6):
        cam = synth.camera(W, H, f)
        chain._set_camera_matrices(f, cam)
        cur64, has = pack(cam)
        if separate:
            # VulkanPBRT.cpp:578-584: prev.view is the push constants' prev_view (identity before the first frame)
            prev64 = np.concatenate([chain.prev_view, np.zeros(48, np.float32)]).astype(np.float32)
        else:
            prev64, _ = pack(prev_cam if prev_cam is not None else cam)
        assert h.hostref_set_camera_matrices(1 if separate else 0, f, p(cur64), has, p(prev64), has if not separate else 0, p(pc)) == 0
        got = np.concatenate([np.array(list(chain.pc.view), np.float32), np.array(list(chain.pc.inv_view), np.float32),
                              np.array(list(chain.pc.prev_view), np.float32), np.array(list(chain.pc.prev_origin), np.float32)])
        assert _same_bits(got, pc[:52]), f"frame {f}"
        assert chain.pc.frame_number == int(pc[52:].view(np.int32)[0]) == f
        chain.prev_view = np.asarray(cam.view, np.float32).copy()        # what run_frame does at the end of a frame
        chain.prev_cam = prev_cam = cam
    # the missing-matrices error of the separate mode (Accumulator.cpp:89-94)
    if separate:
        cur64, _ = pack(synth.camera(W, H, 0))
        assert h.hostref_set_camera_matrices(1, 0, p(cur64), 0, p(cur64), 0, p(pc)) == 1


def test_missing_file_returns_an_empty_list(tmp_path, capsys):
    assert import_matrices(tmp_path / "nope.json") == []
    assert "unable to open" in capsys.readouterr().out


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_offline_mode_driven_from_a_matrix_file_equals_the_oracle(tmp_path, backend, oracle):
    """combined view-projection matrices written to and read from the reference's JSON layout drive the accumulator's
    non-SEPARATE_MATRICES mode; every plane must equal the oracle fed with the same matrices directly"""
    from tests.util import assert_frame_equal, make_pair
    from vulkanpbrt_b200.pipeline import _combined
    W, H, frames = 192, 128, 3
    seq = [synth.render_frame(W, H, f) for f in range(frames)]
    mats = []
    for fr in seq:
        vp, ivp = _combined(fr.camera)
        mats.append(CameraMatrices(view=vp, inv_view=ivp))
    path = tmp_path / "camera.json"
    export_matrices(path, mats)
    loaded = import_matrices(path)
    pipe, orc = make_pair(oracle, W, H, denoiser="bmfr", block=32, use_taa=True, separate_matrices=False)
    for f, fr in enumerate(seq):
        pipe.run_frame_with_matrices(f, fr, loaded)
        pipe.ctx.synchronize()
        orc.run_frame(f, fr)
        assert_frame_equal(pipe, orc, f)


TEXT_FILES = {
    # the BMFR dataset's layout: braces, commas, several matrices, a trailing "};"
    "dataset": lambda cams: "const float camera_matrices[2][4][4] = {\n" + "\n".join(
        "{" + ", ".join(f"{float(x):.9g}" for x in c[:8]) + ",\n " + ", ".join(f"{float(x):.9g}" for x in c[8:]) + "}," for c in cams) + "\n};\n",
    # bare numbers, one per line, no trailing newline after the last one
    "bare": lambda cams: "\n".join(f"{float(x):.9g}" for c in cams for x in c),
    # numbers glued to braces, exponents, a negative zero, a comment word that starts with a digit ("4x4")
    "glued": lambda cams: "matrix 4x4 float\n" + " ".join("{%s}" % f"{float(x):.6e}" for c in cams for x in c) + " -0.0",
}


@pytest.mark.parametrize("name", sorted(TEXT_FILES))
def test_text_matrix_files_import_like_the_reference(tmp_path, name):
    """the BMFR-dataset text branch of MatrixIO::import_matrices (RenderIO.cpp:637-664) as the reference's own text
    (oracle/host_shim) against the Python importer and the C++ one (include/vkpbrt/io.hpp): the same matrices, and the
    same inverses (vsg's), bit for bit"""
    import subprocess
    from pathlib import Path
    h = _hostref()
    cams = [synth.camera(160, 128, f) for f in range(2)]
    mats = []
    for c in cams:
        v = np.asarray(c.view, np.float64).reshape(4, 4).T
        p = np.asarray(c.proj, np.float64).reshape(4, 4).T
        mats.append((p @ v).T.astype(np.float32).reshape(-1))
    path = tmp_path / f"{name}.txt"
    path.write_text(TEXT_FILES[name](mats))
    buf = np.zeros(32 * 8, np.float32)
    n = h.hostref_import_matrices_text(str(path).encode(), buf.ctypes.data_as(C.c_void_p), 8)
    got = import_matrices(path)
    assert len(got) == n and n >= 2
    for i, m in enumerate(got):
        assert _same_bits(np.asarray(m.view, np.float32), buf[32 * i:32 * i + 16]), f"view {i}"
        assert _same_bits(np.asarray(m.inv_view, np.float32), buf[32 * i + 16:32 * i + 32]), f"inverse {i}"
    # the C++ layer: import, then export to JSON, which the Python reader takes back
    root = Path(__file__).resolve().parents[1]
    src = tmp_path / "conv.cpp"
    src.write_text('#include <vkpbrt/io.hpp>\nint main(int, char** a) { auto m = vkpbrt::MatrixIO::import_matrices(a[1]); return vkpbrt::MatrixIO::export_matrices(a[2], m) ? 0 : 1; }\n')
    libdir = root / "tests" / "hostsim"
    subprocess.run(["make", "-C", str(libdir)], check=True, capture_output=True)
    exe = tmp_path / "conv"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(root / "include"), str(src), "-o", str(exe), f"-L{libdir}", "-lvkpbrt_hostsim", f"-Wl,-rpath,{libdir}", "-lz"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(exe), str(path), str(tmp_path / "back.json")]).returncode == 0
    back = import_matrices(tmp_path / "back.json")
    assert len(back) == n
    for i, m in enumerate(back):
        assert _same_bits(np.asarray(m.view, np.float32), buf[32 * i:32 * i + 16]) and _same_bits(np.asarray(m.inv_view, np.float32), buf[32 * i + 16:32 * i + 32])
