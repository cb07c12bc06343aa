"""Random image sizes / wirings / block sizes / modes against the oracle, every plane bit for bit.

Ad-hoc driver behind the bounded, seeded versions in tests/ (tests/test_fuzz_sizes.py); runs on the test emulator
(tests/hostsim), CPU only.  Usage: python tools/fuzz/fuzz_sizes.py <seed> <count>
"""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import random
import numpy as np
from vulkanpbrt_b200 import _capi
_capi._lib = _capi.configure(ctypes.CDLL(str(ROOT / 'tests' / 'hostsim' / 'libvkpbrt_hostsim.so')))
from oracle import oracle as O
from tests.util import make_pair, step_both, assert_frame_equal
seed = int(sys.argv[1]); n = int(sys.argv[2])
rng = random.Random(seed)
fails = 0
for it in range(n):
    den = rng.choice(["bmfr", "bmfr", "bfr", "bfrx3", "bmfrx3"])
    block = rng.choice([8, 16, 32])
    W = rng.choice([1, 2, 3, 5, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 66, 70, 97, 130]) if rng.random() < 0.6 else rng.randint(1, 140)
    H = rng.choice([1, 2, 3, 5, 7, 8, 15, 16, 17, 31, 32, 33, 63, 64, 65, 70, 97]) if rng.random() < 0.6 else rng.randint(1, 140)
    taa = rng.random() < 0.7
    first = rng.choice([0, 0, 6, 7, 8, 14])
    frames = rng.randint(2, 4)
    sep = rng.random() < 0.7
    f16 = rng.random() < 0.2
    pos = rng.choice([0, 0, 0, 1, 2]) if den == "bmfr" else 0
    big = 32 if den.endswith("x3") else block
    W = max(W, big); H = max(H, big)
    if first != 0: sep = True
    cfg = dict(W=W, H=H, den=den, block=block, taa=taa, first=first, frames=frames, sep=sep, f16=f16, pos=pos)
    try:
        pipe, orc = make_pair(O, W, H, denoiser=den, block=block, use_taa=taa, separate_matrices=sep, raw_f16=f16, position_type=pos)
        for f in range(first, first + frames):
            step_both(O, pipe, orc, W, H, f)
            assert_frame_equal(pipe, orc, f)
        print("ok", cfg, flush=True)
    except Exception as e:
        fails += 1
        print("FAIL", cfg, repr(e)[:300], flush=True)
print("fails", fails)
