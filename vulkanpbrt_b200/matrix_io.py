"""Camera-matrix files of the offline mode (the part that needs no EXR reader).

Host-side mirror of `MatrixIO::import_matrices` / `MatrixIO::export_matrices`
(source/io/RenderIO.cpp:593-711): the JSON layout the reference writes and reads
({"amtOfFrames", "matrices": [{"type", "storageType", "view", "invView"[, "proj", "invProj"]}]}, 16 floats per
matrix in vsg's m[col][row] order, i.e. column-major) and the whitespace / comma separated text dump of the BMFR
dataset (one combined view-projection matrix per 16 numbers, inverse computed on load).

The matrices feed `Accumulator.set_camera_matrices` exactly like the reference's `CameraMatricesVec`
(VulkanPBRT.cpp:572-584): with "proj" present the accumulator runs in SEPARATE_MATRICES mode, otherwise in the
combined (offline) mode.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import List

import numpy as np

from .modules import CameraMatrices


def _inverse(m: np.ndarray) -> np.ndarray:
    """4x4 inverse by cofactors in binary32, column-major in and out: the same evaluation order as the library's
    (and the oracle's) mat_inverse, so a matrix loaded from a text file gets the inverse the accumulator would get"""
    f = np.float32
    a = [f(x) for x in m]
    a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33 = a
    b00, b01, b02 = a00 * a11 - a01 * a10, a00 * a12 - a02 * a10, a00 * a13 - a03 * a10
    b03, b04, b05 = a01 * a12 - a02 * a11, a01 * a13 - a03 * a11, a02 * a13 - a03 * a12
    b06, b07, b08 = a20 * a31 - a21 * a30, a20 * a32 - a22 * a30, a20 * a33 - a23 * a30
    b09, b10, b11 = a21 * a32 - a22 * a31, a21 * a33 - a23 * a31, a22 * a33 - a23 * a32
    det = ((((b00 * b11 - b01 * b10) + b02 * b09) + b03 * b08) - b04 * b07) + b05 * b06
    with np.errstate(divide="ignore", invalid="ignore"):
        inv_det = f(1.0) / det
    out = [((a11 * b11 - a12 * b10) + a13 * b09), ((a02 * b10 - a01 * b11) - a03 * b09), ((a31 * b05 - a32 * b04) + a33 * b03),
           ((a22 * b04 - a21 * b05) - a23 * b03), ((a12 * b08 - a10 * b11) - a13 * b07), ((a00 * b11 - a02 * b08) + a03 * b07),
           ((a32 * b02 - a30 * b05) - a33 * b01), ((a20 * b05 - a22 * b02) + a23 * b01), ((a10 * b10 - a11 * b08) + a13 * b06),
           ((a01 * b08 - a00 * b10) - a03 * b06), ((a30 * b04 - a31 * b02) + a33 * b00), ((a21 * b02 - a20 * b04) - a23 * b00),
           ((a11 * b07 - a10 * b09) - a12 * b06), ((a00 * b09 - a01 * b07) + a02 * b06), ((a31 * b01 - a30 * b03) - a32 * b00),
           ((a20 * b03 - a21 * b01) + a22 * b00)]
    return np.array([x * inv_det for x in out], dtype=np.float32)


def _mat(values) -> np.ndarray:
    m = np.asarray(values, dtype=np.float32).reshape(-1)
    if m.size != 16:
        raise ValueError(f"a matrix needs 16 numbers, got {m.size}")
    return m


def import_matrices(matrix_path) -> List[CameraMatrices]:
    """RenderIO.cpp:593-666.  A file that cannot be opened yields an empty list (the reference prints a message and
    returns {}); malformed JSON raises, as nlohmann::json does."""
    path = Path(matrix_path)
    try:
        text = path.read_text()
    except OSError:
        print(f"Matrix file {path} unable to open.")
        return []
    if path.suffix == ".json":
        doc = json.loads(text)
        out = []
        for i in range(int(doc["amtOfFrames"])):
            m = doc["matrices"][i]
            cm = CameraMatrices(view=_mat(m["view"]), inv_view=_mat(m["invView"]))
            if m.get("type") == "ModelView+Projection":
                cm.proj, cm.inv_proj = _mat(m["proj"]), _mat(m["invProj"])
            out.append(cm)
        return out
    # BMFR-dataset text: tokens separated by whitespace, optional trailing ',' and leading '{'; every token that starts
    # with a digit or '-' is a number, every 16 numbers are one matrix (:639-664)
    out, cur = [], []
    for tok in text.split():
        if tok.endswith(","):
            tok = tok[:-1]
        if tok.startswith("{"):
            tok = tok[1:]
        if tok and (tok[0].isdigit() or tok[0] == "-"):
            cur.append(_leading_float(tok))
            if len(cur) == 16:
                m = _mat(cur)
                out.append(CameraMatrices(view=m, inv_view=_inverse(m)))
                cur = []
    return out


def _leading_float(tok: str) -> float:
    """std::stof: the longest valid prefix ("0.5}" -> 0.5)"""
    for end in range(len(tok), 0, -1):
        try:
            return float(tok[:end])
        except ValueError:
            continue
    raise ValueError(f"not a number: {tok!r}")


def export_matrices(matrix_path, matrices: List[CameraMatrices]) -> bool:
    """RenderIO.cpp:668-711 (same keys and values; a matrix without projection is typed "ModelViewProjection")"""
    arr = lambda m: [float(np.float32(x)) for x in np.asarray(m, dtype=np.float32).reshape(-1)]
    objs = []
    for m in matrices:
        has_proj = m.proj is not None and m.inv_proj is not None
        o = {"type": "ModelView+Projection" if has_proj else "ModelViewProjection", "storageType": "ColumnMajor",
             "view": arr(m.view), "invView": arr(m.inv_view)}
        if has_proj:
            o["proj"], o["invProj"] = arr(m.proj), arr(m.inv_proj)
        objs.append(o)
    try:
        Path(matrix_path).write_text(json.dumps({"amtOfFrames": len(matrices), "matrices": objs}, indent=4))
    except OSError:
        print(f"Matrix file {matrix_path} unable to open.")
        return False
    return True
