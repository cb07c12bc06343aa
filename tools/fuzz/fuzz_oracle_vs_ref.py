"""The ORACLE against the reference's own shader source (oracle/_ref, built from /root/reference through oracle/glsl_shim):
random sizes / wirings / modes AND extreme data; every plane bit for bit (NaN payloads canonicalised; the debug feature
buffer / weights only for blocks that hold at least one image pixel -- the reference's single-reflection mirror() reads
out of bounds for the others).

Ad-hoc driver behind tests/test_oracle_vs_ref.py; CPU only, needs /root/reference.  Usage: python tools/fuzz/fuzz_oracle_vs_ref.py <seed> <count>
"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import random
import numpy as np
from vulkanpbrt_b200 import synth
from oracle import oracle as O
from oracle import ref as R
from tests.util import second_moment_plane
R.build()
from tests.test_fuzz_data import _perturb, MODES
seed = int(sys.argv[1]); n = int(sys.argv[2])
rng = random.Random(seed)
def canon(a):
    a = np.asarray(a)
    if a.dtype == np.uint16:
        a = a.copy(); a[(a & 0x7FFF) > 0x7C00] = 0x7E00
    if a.dtype == np.float32:
        a = a.copy(); a[np.isnan(a)] = np.float32(np.nan)
        return a.view(np.uint32)
    return a
fails = 0
for it in range(n):
    den = rng.choice(["bmfr", "bmfr", "bfr", "bfrx3", "bmfrx3"]); block = rng.choice([8, 16, 32])
    big = 32 if den.endswith("x3") else block
    W = max(big, rng.choice([1, 8, 17, 31, 32, 33, 40, 54, 63, 64, 65, 70, 97]) if rng.random() < 0.6 else rng.randint(1, 110))
    H = max(big, rng.choice([1, 8, 17, 31, 32, 33, 40, 63, 64, 65, 70]) if rng.random() < 0.6 else rng.randint(1, 110))
    taa = rng.random() < 0.7; first = rng.choice([0, 0, 7, 14]); frames = rng.randint(2, 3)
    sep = True if first else rng.random() < 0.7; f16 = rng.random() < 0.2
    pos = rng.choice([0, 0, 1, 2]) if den == "bmfr" and block in (16, 32) else 0
    modes = [rng.choice(MODES + ["none", "none"]) for _ in range(frames)]
    cfg = dict(W=W, H=H, den=den, block=block, taa=taa, first=first, frames=frames, sep=sep, f16=f16, pos=pos, modes=modes)
    nrng = np.random.default_rng(seed * 1000 + it)
    try:
        a = O.OracleChain(W, H, den, block, use_taa=taa, separate_matrices=sep, raw_f16=f16, position_type=pos)
        b = R.RefChain(W, H, den, block, use_taa=taa, separate_matrices=sep, raw_f16=f16, position_type=pos)
        for k, f in enumerate(range(first, first + frames)):
            fr = _perturb(synth.render_frame(W, H, f), nrng, modes[k])
            if den.endswith("x3"):
                sq = second_moment_plane(O, a); a.average_squared[...] = sq; b.average_squared[...] = sq
            a.run_frame(f, fr, keep_debug=True); b.run_frame(f, fr, keep_debug=True)
            for name in ("motion", "spp", "illum", "prev_depth", "blend_final", "taa_final", "taa_history"):
                assert np.array_equal(canon(getattr(a, name)), canon(getattr(b, name))), (name, f)
            for blk in a.blocks:
                assert np.array_equal(canon(a.denoised[blk]), canon(b.denoised[blk])), ("denoised", blk, f)
                assert np.array_equal(a.finals[blk], b.finals[blk]), ("final", blk, f)
            if den == "bmfr":
                ox, oy = O.bmfr_block_offset(block, f)
                Hb, Wb = a.weights.shape[1:]
                inside = np.zeros((Hb, Wb), bool)
                for by in range(Hb):
                    for bx in range(Wb):
                        x0, y0 = bx * block - ox, by * block - oy
                        inside[by, bx] = x0 < W and x0 + block > 0 and y0 < H and y0 + block > 0
                wa, wb = canon(a.weights), canon(b.weights)
                assert np.array_equal(wa[:, inside], wb[:, inside]), ("weights", f)
                big = np.repeat(np.repeat(inside, block, 0), block, 1)
                fa, fb = canon(a.features), canon(b.features)
                for x in (fa, fb): x[x == 0x8000] = 0      # -0 / +0 depths in one block: which zero a min / max reduction returns depends on its order
                assert np.array_equal(fa[:, big], fb[:, big]), ("features", f)
        print("ok", cfg, flush=True)
    except AssertionError as e:
        fails += 1; print("FAIL", cfg, e, flush=True)
    except Exception as e:
        fails += 1; print("ERR", cfg, repr(e)[:200], flush=True)
print("fails", fails)
