"""Generates the committed golden hashes under tests/golden/.

The reference cannot be executed in this environment (GLSL; no glslc / Vulkan ICD) and ships no test
vectors, so these goldens come from the oracle (oracle/vkpbrt_oracle.c) on the deterministic synthetic
sequence.  They pin the oracle bit for bit; the CUDA path is compared with the oracle directly.

    python -m tests.golden.make_golden        # rewrites tests/golden/*.json
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from vulkanpbrt_b200 import synth  # noqa: E402

CASES = {
    "bmfr32_taa_256x256_8f": dict(W=256, H=256, denoiser="bmfr", block=32, taa=True, frames=8),
    "bfrx3_taa_160x128_3f": dict(W=160, H=128, denoiser="bfrx3", block=32, taa=True, frames=3),
}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(oracle, c):
    orc = oracle.OracleChain(c["W"], c["H"], c["denoiser"], c["block"], use_taa=c["taa"])
    frames = []
    for f in range(c["frames"]):
        fr = synth.render_frame(c["W"], c["H"], f)
        if c["denoiser"].endswith("x3"):
            av = oracle.f16_bits_to_f32(orc.prev_illu)
            orc.average_squared[...] = (av * av * 1.5 + 0.01).astype(np.float16).view(np.uint16)
        orc.run_frame(f, fr)
        frames.append({
            "input": _sha(np.concatenate([fr.depth.view(np.uint8).ravel(), fr.normal.view(np.uint8).ravel(),
                                          fr.albedo.ravel(), fr.illumination.view(np.uint8).ravel()])),
            "motion": _sha(orc.motion), "spp": _sha(orc.spp), "illum": _sha(orc.illum),
            "denoised": {str(b): _sha(orc.denoised[b]) for b in orc.blocks},
            "denoiser_final": _sha(orc.denoiser_final()), "final": _sha(orc.final()),
        })
    return {"case": c, "frames": frames}


if __name__ == "__main__":
    from oracle import oracle as O
    out = Path(__file__).resolve().parent
    for name, c in CASES.items():
        (out / f"{name}.json").write_text(json.dumps(run_case(O, c), indent=1))
        print("wrote", name)
