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
        L.vkpbrt_oracle_mat_inverse((C.c_float * 16)(*[float(x) for x in c.view]), want)
        np.testing.assert_array_equal(m.inv_view.view(np.uint32), np.array(list(want), np.float32).view(np.uint32))
        assert m.proj is None


def test_inverse_matches_the_oracle_bit_for_bit(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(7)
    for _ in range(50):
        m = rng.standard_normal(16).astype(np.float32)
        want = (C.c_float * 16)()
        L.vkpbrt_oracle_mat_inverse((C.c_float * 16)(*[float(x) for x in m]), want)
        np.testing.assert_array_equal(_inverse(m).view(np.uint32), np.array(list(want), np.float32).view(np.uint32))


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
