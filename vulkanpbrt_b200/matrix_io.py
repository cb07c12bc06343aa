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
    """vsg::inverse(mat4), the routine the reference's matrix import calls (RenderIO.cpp:659): the library's restatement
    (vkpbrt_mat4_inverse -- host code, no device needed), so a matrix loaded from a text file gets the inverse bits the
    reference would compute"""
    from . import _capi as capi
    a = np.ascontiguousarray(m, dtype=np.float32).reshape(16)
    out = np.empty(16, np.float32)
    capi.call("vkpbrt_mat4_inverse", a.ctypes.data, out.ctypes.data)
    return out


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
