"""Seeded fuzz over image sizes and configurations on the emulator: ragged and minimal sizes (down to one block),
every denoiser wiring, block size, position mode, matrix mode, input format, starting frame -- every plane of every
frame against the oracle, bit for bit.

Why it exists: the first run of this fuzz found that images narrower than two 32-blocks used a partly stale
block-invariant table from frame 1 on (the launch's spare CTAs were too few to build it) -- a case none of the
hand-picked sizes covered.  The configurations are drawn from a fixed seed, so the test is deterministic; it runs on
the emulator only (the GPU suite keeps its hand-picked, GPU-confirmed sizes)."""
import random

import pytest

from tests.conftest import backend_params
from tests.util import assert_frame_equal, make_pair, step_both

SIZES = [1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 40, 54, 63, 64, 65, 66, 70, 97, 130]


def _configs(seed, n):
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        den = rng.choice(["bmfr", "bmfr", "bfr", "bfrx3", "bmfrx3"])
        block = rng.choice([8, 16, 32])
        W = rng.choice(SIZES) if rng.random() < 0.6 else rng.randint(1, 140)
        H = rng.choice(SIZES) if rng.random() < 0.6 else rng.randint(1, 140)
        big = 32 if den.endswith("x3") else block
        W, H = max(W, big), max(H, big)                    # smaller than one block is rejected (see test below)
        first = rng.choice([0, 0, 6, 7, 8, 14])            # 8, 9: negative jitters; 15 -> 16: the jitter table wraps
        out.append(dict(W=W, H=H, den=den, block=block, taa=rng.random() < 0.7, first=first, frames=rng.randint(2, 3),
                        sep=True if first else rng.random() < 0.7, f16=rng.random() < 0.2, pos=rng.choice([0, 0, 0, 1, 2]) if den == "bmfr" else 0))
    return out


@pytest.mark.parametrize("backend", [backend_params()[0]], indirect=True)
@pytest.mark.parametrize("cfg", _configs(20261017, 40), ids=lambda c: f"{c['den']}{c['block']}-{c['W']}x{c['H']}-f{c['first']}")
def test_random_sizes_and_configurations(backend, oracle, cfg):
    W, H = cfg["W"], cfg["H"]
    pipe, orc = make_pair(oracle, W, H, denoiser=cfg["den"], block=cfg["block"], use_taa=cfg["taa"], separate_matrices=cfg["sep"],
                          raw_f16=cfg["f16"], position_type=cfg["pos"])
    for f in range(cfg["first"], cfg["first"] + cfg["frames"]):
        step_both(oracle, pipe, orc, W, H, f)
        assert_frame_equal(pipe, orc, f)


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_images_smaller_than_one_block_are_rejected(backend):
    """the reference's mirror() reflects once (bmfrGeneral.comp:93-101): with an image smaller than the block a mirrored
    coordinate can still lie outside it and the shaders read out of bounds; the replacement refuses such sizes"""
    from vulkanpbrt_b200 import BFR, BMFR, Accumulator, Context, GBuffer, IlluminationBufferDemodulatedFloat, VkpbrtError, _capi
    ctx = Context(0)
    for (w, h) in ((31, 64), (64, 31)):
        g, raw = GBuffer.create(ctx, w, h), IlluminationBufferDemodulatedFloat.create(ctx, w, h)
        acc = Accumulator.create(g, raw, True)
        for cls in (BMFR, BFR):
            with pytest.raises(VkpbrtError) as e:
                cls.create(w, h, 32, 32, g, acc.accumulated_illumination, acc.accumulation_buffer)
            assert e.value.code == _capi.ERR_INVALID_ARGUMENT and "smaller than one block" in str(e.value)
            cls.create(w, h, 16, 16, g, acc.accumulated_illumination, acc.accumulation_buffer)      # fine with a smaller block


@pytest.mark.parametrize("backend", [backend_params()[0]], indirect=True)
@pytest.mark.parametrize("den,block,start", [("bmfr", 32, 2**32 - 3), ("bmfr", 8, 2**31 - 2), ("bfr", 16, 2**32 - 3), ("bmfr", 16, 65534)])
def test_frame_numbers_up_to_the_uint32_wrap(backend, oracle, den, block, start):
    """frameNumber is a uint (PipelineStructs.hpp:12): the noise seed (bmfrGeneral.comp:115-116), the jitter phase, the
    ping-pong layer and the per-frame table's frame + 1 all wrap with it; 2^32 - 1 is followed by frame 0 (no history)"""
    from vulkanpbrt_b200 import synth
    W, H = 64, 40
    pipe, orc = make_pair(oracle, W, H, denoiser=den, block=block, use_taa=True)
    for k in range(5):
        f = (start + k) % 2**32
        fr = synth.render_frame(W, H, k)
        pipe.run_frame(f, fr)
        pipe.ctx.synchronize()
        orc.run_frame(f, fr)
        assert_frame_equal(pipe, orc, f)
