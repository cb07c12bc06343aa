"""Extreme input DATA (scaled illumination, special values, noise / flat G-buffers, camera jumps) against the oracle.

Ad-hoc driver behind the bounded, seeded versions in tests/ (tests/test_fuzz_data.py); runs on the test emulator
(tests/hostsim), CPU only.  Usage: python tools/fuzz/fuzz_data.py <seed> <count> [nan]   (nan: also NaN inputs -- outside the parity contract)
"""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import random
import numpy as np
from vulkanpbrt_b200 import _capi, synth
_capi._lib = _capi.configure(ctypes.CDLL(str(ROOT / 'tests' / 'hostsim' / 'libvkpbrt_hostsim.so')))
from oracle import oracle as O
from tests.util import make_pair
import tests.util as U
def canon(a):
    a = np.asarray(a)
    if a.dtype == np.uint16:
        a = a.copy(); a[(a & 0x7FFF) > 0x7C00] = 0x7E00
    return a
_orig = np.testing.assert_array_equal
def _cmp(a, b, err_msg=""):
    _orig(canon(a), canon(b), err_msg=err_msg)
def assert_frame_equal(pipe, orc, f):
    np.testing.assert_array_equal = _cmp
    try: U.assert_frame_equal(pipe, orc, f)
    finally: np.testing.assert_array_equal = _orig
seed = int(sys.argv[1]); n = int(sys.argv[2]); allow_nan = len(sys.argv) > 3
rng = random.Random(seed)
def perturb(fr, nrng, mode):
    H, W = fr.depth.shape
    if mode == "scale":
        s = np.float32(10.0 ** nrng.uniform(-38, 38)); fr.illumination[...] = (fr.illumination.astype(np.float64) * float(s)).astype(np.float32)
    elif mode == "specials":
        vals = np.array([0.0, -0.0, 1e-45, 1e-38, 1e-20, 1e20, 3e38, np.inf, -1.0, -np.inf] + ([np.nan] if allow_nan else []), np.float32)
        m = nrng.random((H, W)) < 0.05
        fr.illumination[m] = nrng.choice(vals, size=(int(m.sum()), 4))
        m = nrng.random((H, W)) < 0.05
        fr.depth[m] = nrng.choice(vals, size=int(m.sum()))
        m = nrng.random((H, W)) < 0.05
        fr.normal[m] = nrng.choice(np.array([0, 1e-30, 3.14159274, 6.5, -7.1, 1e6, 1e20, 3e38], np.float32), size=(int(m.sum()), 2))
    elif mode == "noise":
        fr.depth[...] = (10.0 ** nrng.uniform(-3, 6, (H, W))).astype(np.float32)
        fr.normal[...] = nrng.uniform(-10, 10, (H, W, 2)).astype(np.float32)
        fr.albedo[...] = nrng.integers(0, 256, (H, W, 4), dtype=np.uint8)
        fr.illumination[...] = (10.0 ** nrng.uniform(-10, 4, (H, W, 4))).astype(np.float32)
    elif mode == "flat":
        fr.depth[...] = np.float32(nrng.choice([1.0, 7.25, 1e10]))
        fr.normal[...] = np.float32(nrng.choice([0.0, 0.7853982]))
        fr.illumination[...] = np.float32(nrng.choice([0.0, 0.5, 10.0]))
    elif mode == "camera":
        R = np.eye(4); a = nrng.uniform(-0.5, 0.5)
        R[:2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]; R[:3, 3] = nrng.uniform(-2, 2, 3)
        v = fr.camera.view.reshape(4, 4).T.astype(np.float64) @ R
        fr.camera.view = v.T.astype(np.float32).reshape(-1).copy(); fr.camera.inv_view = np.linalg.inv(v).T.astype(np.float32).reshape(-1).copy()
    return fr
fails = 0
for it in range(n):
    den = rng.choice(["bmfr", "bmfr", "bfr"]); block = rng.choice([8, 16, 32])
    W = rng.choice([32, 40, 64, 66, 97]); H = rng.choice([32, 40, 64, 70])
    taa = rng.random() < 0.6; first = rng.choice([0, 7]); frames = 3
    modes = [rng.choice(["scale", "specials", "noise", "flat", "camera", "none"]) for _ in range(frames)]
    cfg = dict(W=W, H=H, den=den, block=block, taa=taa, first=first, modes=modes, seed=(seed, it))
    nrng = np.random.default_rng(seed * 1000 + it)
    try:
        pipe, orc = make_pair(O, W, H, denoiser=den, block=block, use_taa=taa)
        for k, f in enumerate(range(first, first + frames)):
            fr = perturb(synth.render_frame(W, H, f), nrng, modes[k])
            pipe.run_frame(f, fr); pipe.ctx.synchronize(); orc.run_frame(f, fr)
            assert_frame_equal(pipe, orc, f)
        print("ok", cfg, flush=True)
    except Exception as e:
        fails += 1
        print("FAIL", cfg, repr(e)[:300].replace("\n", " "), flush=True)
print("fails", fails)
