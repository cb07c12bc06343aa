"""Deterministic synthetic G-buffer sequence; ctypes binding of
vulkanpbrt_b200/synth/synth.c.  Input tooling for tests, smoke and bench -- not part of the
denoising path."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Tuple

import numpy as np

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libvkpbrt_synth.so"
SEED = 0x5EED0001


class _Camera(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("inv_view", C.c_float * 16), ("proj", C.c_float * 16), ("inv_proj", C.c_float * 16)]


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise ImportError(f"{_LIB_PATH} is missing: run `python -m vulkanpbrt_b200.build`")
        l = C.CDLL(str(_LIB_PATH))
        l.vkpbrt_synth_camera.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(_Camera)]
        l.vkpbrt_synth_camera.restype = None
        l.vkpbrt_synth_frame.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int] + [C.c_void_p] * 5
        l.vkpbrt_synth_frame.restype = None
        _lib = l
    return _lib


@dataclass
class Camera:
    view: np.ndarray       # float32[16], column-major
    inv_view: np.ndarray
    proj: np.ndarray
    inv_proj: np.ndarray


def camera(width: int, height: int, frame: int) -> Camera:
    c = _Camera()
    _load().vkpbrt_synth_camera(width, height, frame, C.byref(c))
    f = lambda a: np.array(list(a), dtype=np.float32)
    return Camera(f(c.view), f(c.inv_view), f(c.proj), f(c.inv_proj))


@dataclass
class Frame:
    index: int
    depth: np.ndarray          # float32 [H][W]
    normal: np.ndarray         # float32 [H][W][2]  (theta, phi)
    albedo: np.ndarray         # uint8   [H][W][4]  RGBA
    material: np.ndarray       # uint8   [H][W][4]
    illumination: np.ndarray   # float32 [H][W][4]  demodulated 1-spp radiance
    camera: Camera

    @property
    def nbytes(self) -> int:
        return self.depth.nbytes + self.normal.nbytes + self.albedo.nbytes + self.illumination.nbytes


def render_frame(width: int, height: int, frame: int, seed: int = SEED, rows: Optional[Tuple[int, int]] = None,
                 out: Optional[Frame] = None, alloc=np.empty) -> Frame:
    """Renders frame `frame` (all rows, or image rows [rows[0], rows[1]) into full-size planes)."""
    if out is None:
        z = np.zeros if rows is not None else alloc
        out = Frame(frame, z((height, width), np.float32), z((height, width, 2), np.float32),
                    z((height, width, 4), np.uint8), z((height, width, 4), np.uint8),
                    z((height, width, 4), np.float32), camera(width, height, frame))
    else:
        out.index = frame
        out.camera = camera(width, height, frame)
    r0, r1 = rows if rows is not None else (0, height)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    _load().vkpbrt_synth_frame(width, height, frame, seed & 0xFFFFFFFF, r0, r1, p(out.depth), p(out.normal),
                               p(out.albedo), p(out.material), p(out.illumination))
    return out
