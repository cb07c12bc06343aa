"""Seeded fuzz over the DATA on the emulator: illumination scaled across the whole binary32 range, special values
(zeros, denormals, huge values, infinities) sprinkled over illumination / depth / normals, white-noise G-buffers,
perfectly flat G-buffers (the fit's rank-deficient case), camera jumps far beyond the usual frame-to-frame motion.
These drive the kernels' guarded fast paths (exact reciprocal division, in-range sqrt / reciprocal, the generic QR
fallback) into their out-of-range branches.  Every plane must equal the oracle bit for bit -- except the payload and
sign of NaNs, which IEEE 754 leaves open (x86 produces the negative default NaN, NVIDIA GPUs 0x7fffffff): fp16 NaNs
are canonicalised before comparing.  Emulator only (see tests/test_fuzz_sizes.py)."""
import random

import numpy as np
import pytest

import tests.util as U
from tests.conftest import backend_params
from vulkanpbrt_b200 import synth

MODES = ["scale", "specials", "noise", "flat", "camera", "none"]


def _perturb(fr, rng, mode):
    H, W = fr.depth.shape
    if mode == "scale":
        fr.illumination[...] = (fr.illumination.astype(np.float64) * 10.0 ** rng.uniform(-38, 38)).astype(np.float32)
    elif mode == "specials":
        vals = np.array([0.0, -0.0, 1e-45, 1e-38, 1e-20, 1e20, 3e38, np.inf, -1.0, -np.inf], np.float32)
        m = rng.random((H, W)) < 0.05
        fr.illumination[m] = rng.choice(vals, size=(int(m.sum()), 4))
        m = rng.random((H, W)) < 0.05
        fr.depth[m] = rng.choice(vals, size=int(m.sum()))
        m = rng.random((H, W)) < 0.05
        fr.normal[m] = rng.choice(np.array([0, 1e-30, 3.14159274, 6.5, -7.1, 1e6, 1e20, 3e38], np.float32), size=(int(m.sum()), 2))
    elif mode == "noise":
        fr.depth[...] = (10.0 ** rng.uniform(-3, 6, (H, W))).astype(np.float32)
        fr.normal[...] = rng.uniform(-10, 10, (H, W, 2)).astype(np.float32)
        fr.albedo[...] = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        fr.illumination[...] = (10.0 ** rng.uniform(-10, 4, (H, W, 4))).astype(np.float32)
    elif mode == "flat":
        fr.depth[...] = np.float32(rng.choice([1.0, 7.25, 1e10]))
        fr.normal[...] = np.float32(rng.choice([0.0, 0.7853982]))
        fr.illumination[...] = np.float32(rng.choice([0.0, 0.5, 10.0]))
    elif mode == "camera":
        R = np.eye(4)
        a = rng.uniform(-0.5, 0.5)
        R[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
        R[:3, 3] = rng.uniform(-2, 2, 3)
        v = fr.camera.view.reshape(4, 4).T.astype(np.float64) @ R
        fr.camera.view = v.T.astype(np.float32).reshape(-1).copy()
        fr.camera.inv_view = np.linalg.inv(v).T.astype(np.float32).reshape(-1).copy()
    return fr


def _configs(seed, n):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        out.append(dict(den=rng.choice(["bmfr", "bmfr", "bfr"]), block=rng.choice([8, 16, 32]), W=rng.choice([32, 40, 64, 66, 97]),
                        H=rng.choice([32, 40, 64, 70]), taa=rng.random() < 0.6, first=rng.choice([0, 7]),
                        modes=[rng.choice(MODES) for _ in range(3)], seed=seed * 1000 + i))
    return out


def _canon(a):
    a = np.asarray(a)
    if a.dtype == np.uint16:                        # fp16 bit patterns: any NaN -> one NaN
        a = a.copy()
        a[(a & 0x7FFF) > 0x7C00] = 0x7E00
    return a


@pytest.mark.parametrize("backend", [backend_params()[0]], indirect=True)
@pytest.mark.parametrize("cfg", _configs(7, 24), ids=lambda c: f"{c['den']}{c['block']}-{c['W']}x{c['H']}-{'-'.join(c['modes'])}")
def test_extreme_inputs(backend, oracle, cfg, monkeypatch):
    exact = np.testing.assert_array_equal
    monkeypatch.setattr(np.testing, "assert_array_equal", lambda a, b, err_msg="": exact(_canon(a), _canon(b), err_msg=err_msg))
    W, H = cfg["W"], cfg["H"]
    rng = np.random.default_rng(cfg["seed"])
    pipe, orc = U.make_pair(oracle, W, H, denoiser=cfg["den"], block=cfg["block"], use_taa=cfg["taa"])
    for k, f in enumerate(range(cfg["first"], cfg["first"] + 3)):
        fr = _perturb(synth.render_frame(W, H, f), rng, cfg["modes"][k])
        pipe.run_frame(f, fr)
        pipe.ctx.synchronize()
        orc.run_frame(f, fr)
        U.assert_frame_equal(pipe, orc, f)
