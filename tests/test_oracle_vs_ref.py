"""Pins the oracle: oracle/vkpbrt_oracle.c (the hand-written restatement every parity test compares the CUDA path
with) against oracle/_ref -- the reference's OWN shader source text (shaders/*.comp under /root/reference), compiled
as C++ through oracle/glsl_shim and executed on the CPU with the reference's descriptor bindings, dispatch sizes and
push constants (oracle/ref.py).  Every plane of every frame must be bit-identical, including the BMFR feature buffer
and the fitted weights.  oracle/_ref is built here (where the reference is mounted) by __graft_entry__.build(); the
prebuilt library travels to the GPU box."""
import numpy as np
import pytest

from vulkanpbrt_b200 import synth


@pytest.fixture(scope="module")
def ref(oracle):
    from oracle import ref as R
    try:
        ok = R.build()
    except Exception as e:      # pragma: no cover
        pytest.fail(f"oracle/_ref failed to build: {e}")
    if not ok:
        pytest.skip("oracle/_ref is not built and /root/reference is not mounted: parity unpinned on this machine")
    return R


CASES = [
    # W, H, denoiser, block, taa, frames, first frame, separate matrices, rgba16f input
    (256, 256, "bmfr", 32, True, 4, 0, True, False),        # BASELINE configs[0] geometry
    (250, 130, "bmfr", 32, True, 4, 7, True, False),        # ragged size, frames 8 / 9 leave rows 0-1 / column 0 unwritten
    (208, 144, "bmfr", 16, False, 2, 8, True, False),
    (200, 136, "bmfr", 8, True, 2, 0, True, False),
    (160, 128, "bfr", 32, False, 2, 0, True, False),
    (168, 104, "bfr", 16, False, 2, 0, True, False),
    (128, 96, "bfr", 8, False, 2, 0, True, False),
    (160, 128, "bfrx3", 32, True, 3, 0, True, False),       # BASELINE configs[2] structure: 3 x BFR + blender (+ TAA)
    (160, 128, "bmfrx3", 32, True, 3, 7, True, False),      # 3 x BMFR + blender (DenoiserUtils.cpp:106-124), frames 7-9
    (256, 128, "bmfr", 32, True, 3, 0, False, False),       # accumulator.comp without SEPARATE_MATRICES
    (256, 128, "bmfr", 32, False, 2, 0, True, True),        # rgba16f raw illumination
]


@pytest.mark.parametrize("W,H,den,block,taa,frames,f0,sep,f16", CASES)
def test_oracle_equals_reference_shader_source(oracle, ref, W, H, den, block, taa, frames, f0, sep, f16):
    a = oracle.OracleChain(W, H, den, block, use_taa=taa, separate_matrices=sep, raw_f16=f16)
    b = ref.RefChain(W, H, den, block, use_taa=taa, separate_matrices=sep, raw_f16=f16)
    for f in range(f0, f0 + frames):
        fr = synth.render_frame(W, H, f)
        if den.endswith("x3"):
            av = oracle.f16_bits_to_f32(a.prev_illu)
            sq = np.ascontiguousarray((av * av * 1.5 + 0.01).astype(np.float16).view(np.uint16))
            a.average_squared[...] = sq
            b.average_squared[...] = sq
        a.run_frame(f, fr, keep_debug=True)
        b.run_frame(f, fr, keep_debug=True)
        for name in ("motion", "spp", "illum", "prev_depth", "blend_final", "taa_final", "taa_history"):
            np.testing.assert_array_equal(getattr(a, name), getattr(b, name), err_msg=f"{name}, frame {f}")
        for blk in a.blocks:
            np.testing.assert_array_equal(a.denoised[blk], b.denoised[blk], err_msg=f"denoised b={blk}, frame {f}")
            np.testing.assert_array_equal(a.finals[blk], b.finals[blk], err_msg=f"final b={blk}, frame {f}")
        if den.startswith("bmfr"):
            np.testing.assert_array_equal(a.features, b.features, err_msg=f"feature buffer, frame {f}")
            np.testing.assert_array_equal(a.weights.view(np.uint32), b.weights.view(np.uint32), err_msg=f"weights, frame {f}")


@pytest.mark.parametrize("ptype,block,W,H", [(1, 32, 160, 128), (2, 32, 160, 128), (1, 16, 112, 80), (2, 16, 112, 80)])
def test_world_position_modes(oracle, ref, ptype, block, W, H):
    """bmfrPre.comp:45-76 / bmfrPost.comp:40-71 with POSITION_TYPE = 1 (POSITION_WORLD_DEPTH_NORM) and 2 (POSITION_WORLD):
    the feature buffer now depends on the camera matrices of the push constants; frames 7-9 cover both negative jitters"""
    a = oracle.OracleChain(W, H, "bmfr", block, use_taa=False, position_type=ptype)
    b = ref.RefChain(W, H, "bmfr", block, use_taa=False, position_type=ptype)
    for f in range(7, 10):
        fr = synth.render_frame(W, H, f)
        a.run_frame(f, fr, keep_debug=True)
        b.run_frame(f, fr, keep_debug=True)
        np.testing.assert_array_equal(a.features, b.features, err_msg=f"feature buffer, frame {f}")
        np.testing.assert_array_equal(a.weights.view(np.uint32), b.weights.view(np.uint32), err_msg=f"weights, frame {f}")
        np.testing.assert_array_equal(a.denoised[block], b.denoised[block], err_msg=f"denoised, frame {f}")
        np.testing.assert_array_equal(a.finals[block], b.finals[block], err_msg=f"final, frame {f}")
    # the modes are not degenerate copies of POSITION_DEPTH
    c = oracle.OracleChain(W, H, "bmfr", block, use_taa=False)
    c.run_frame(7, synth.render_frame(W, H, 7), keep_debug=True)
    a2 = oracle.OracleChain(W, H, "bmfr", block, use_taa=False, position_type=ptype)
    a2.run_frame(7, synth.render_frame(W, H, 7), keep_debug=True)
    assert not np.array_equal(a2.features[4:7], c.features[4:7])


@pytest.mark.parametrize("den,block", [("bmfr", 32), ("bfr", 16)])
def test_history_sampling_with_motion_outside_the_unit_square(oracle, ref, den, block, monkeypatch):
    """texture() with REPEAT addressing at uv beyond [0,1] (a caller-supplied motion plane): the oracle's sampler against
    the shim's, through bmfrPost.comp:108-113 / bfr.comp:293-298; the CUDA side of this case is
    tests/test_parity.py::test_caller_supplied_motion_outside_the_unit_square"""
    import copy
    W, H = 96, 72
    a = oracle.OracleChain(W, H, den, block, use_taa=True)
    b = ref.RefChain(W, H, den, block, use_taa=True)
    rng = np.random.default_rng(17)
    real_dispatch = ref.dispatch
    for f in range(3):
        fr = synth.render_frame(W, H, f)
        probe = copy.deepcopy(a)
        probe.run_frame(f, fr)
        motion = probe.motion.copy()                       # what the accumulator writes this frame ...
        if f > 0:                                          # ... with a third of the vectors replaced
            uv = rng.choice(np.array([1.0, 1.25, 2.5, 37.75, 1000.5, 0.0, 0.999, 3.0], np.float16), size=(H, W, 2))
            pick = rng.random((H, W)) < 0.3
            motion[pick] = uv.view(np.uint16)[pick]
        injected = []

        def dispatch(shader, *args, **kw):
            if not shader.startswith("accumulator") and not injected:
                b.motion[...] = motion
                injected.append(shader)
            return real_dispatch(shader, *args, **kw)

        monkeypatch.setattr(ref, "dispatch", dispatch)
        a.run_frame(f, fr, motion_override=motion)
        b.run_frame(f, fr)
        monkeypatch.setattr(ref, "dispatch", real_dispatch)
        assert injected
        for name in ("motion", "spp", "illum", "taa_final"):
            np.testing.assert_array_equal(getattr(a, name), getattr(b, name), err_msg=f"{name}, frame {f}")
        np.testing.assert_array_equal(a.denoised[block], b.denoised[block], err_msg=f"denoised, frame {f}")
        np.testing.assert_array_equal(a.finals[block], b.finals[block], err_msg=f"final, frame {f}")


def _canon_nan(a):
    a = np.asarray(a)
    if a.dtype == np.uint16:            # fp16 bits: NaN payload / sign are not defined by IEEE 754
        a = a.copy()
        a[(a & 0x7FFF) > 0x7C00] = 0x7E00
    return a


@pytest.mark.parametrize("block,W,H,px", [(8, 64, 48, (7, 0)), (16, 64, 48, (7, 0)), (32, 96, 64, (10, 34))])
def test_bfr_descent_stops_on_the_gradients_the_shader_actually_sums(oracle, ref, block, W, H, px):
    """bfr.comp:260 leaves the descent when gradient_rest turns NaN.  gradient_left (:136-137) is a subgroupAdd over the
    threads with id < 7 that share thread 0's SUBGROUP, and ID = x * size.y + y (:75) spreads those threads over
    subgroups of 32 / b features: only features j < 32 / b (4, 2, 1) can end the loop.  An infinite history value at a sky
    pixel next to the horizon (normal (0, 0, 1): features 4 and 5 are exactly 0, 0 * inf = NaN there and +-inf in the
    others) separates the two readings: summing all seven gradients stops after one iteration, the shader runs on and
    the FINITE channels of that block end up different.  Found by fuzzing the oracle against the shader source."""
    a = oracle.OracleChain(W, H, "bfr", block)
    b = ref.RefChain(W, H, "bfr", block)
    fr7, fr8 = synth.render_frame(W, H, 7), synth.render_frame(W, H, 8)
    assert fr8.depth[px] > 1e9                      # a sky pixel
    a.run_frame(7, fr7)
    b.run_frame(7, fr7)
    for ch in (a, b):
        ch.prev_illu[px[0], px[1], 0] = 0x7C00      # +inf in the red channel of the accumulated history
    a.run_frame(8, fr8)
    b.run_frame(8, fr8)
    assert a.illum[px[0], px[1], 0] == 0x7C00        # it reached the denoiser's input
    for name in ("motion", "spp", "illum"):
        np.testing.assert_array_equal(_canon_nan(getattr(a, name)), _canon_nan(getattr(b, name)), err_msg=name)
    np.testing.assert_array_equal(_canon_nan(a.denoised[block]), _canon_nan(b.denoised[block]), err_msg="denoised")
    np.testing.assert_array_equal(a.finals[block], b.finals[block], err_msg="final")
    red_nan = (a.denoised[block][1, ..., 0] & 0x7FFF) > 0x7C00
    assert red_nan.any() and not ((a.denoised[block][1, ..., 1] & 0x7FFF) > 0x7C00).any()     # red is lost in that block, green is not


def _fuzz_configs(seed, n):
    import random
    from tests.test_fuzz_data import MODES
    rng = random.Random(seed)
    out = []
    for i in range(n):
        den = rng.choice(["bmfr", "bmfr", "bfr", "bfrx3", "bmfrx3"])
        block = rng.choice([8, 16, 32])
        big = 32 if den.endswith("x3") else block
        W = max(big, rng.choice([1, 8, 17, 31, 32, 33, 40, 54, 63, 64, 65, 70, 97]) if rng.random() < 0.6 else rng.randint(1, 110))
        H = max(big, rng.choice([1, 8, 17, 31, 32, 33, 40, 63, 64, 65, 70]) if rng.random() < 0.6 else rng.randint(1, 110))
        first = rng.choice([0, 0, 7, 14])
        frames = rng.randint(2, 3)
        out.append(dict(W=W, H=H, den=den, block=block, taa=rng.random() < 0.7, first=first, frames=frames, sep=True if first else rng.random() < 0.7,
                        f16=rng.random() < 0.2, pos=rng.choice([0, 0, 1, 2]) if den == "bmfr" and block in (16, 32) else 0,
                        modes=[rng.choice(MODES + ["none", "none"]) for _ in range(frames)], seed=seed * 1000 + i))
    return out


@pytest.mark.parametrize("cfg", _fuzz_configs(5, 14), ids=lambda c: f"{c['den']}{c['block']}-{c['W']}x{c['H']}-f{c['first']}-{'-'.join(c['modes'])}")
def test_seeded_fuzz_against_the_shader_source(oracle, ref, cfg):
    """the oracle against the reference's shader source over random sizes (down to one block), wirings, modes AND extreme
    data (tests/test_fuzz_data.py's perturbations): every plane, NaN payloads canonicalised; feature buffer and weights
    for the blocks that hold at least one image pixel (for the others the reference's single-reflection mirror() reads
    out of bounds and the two sides differ, with no effect on any output).  tools/fuzz/fuzz_oracle_vs_ref.py is the
    unbounded version; it found the descent-exit discrepancy fixed above."""
    from tests.test_fuzz_data import _perturb
    from tests.util import second_moment_plane
    W, H, den, block = cfg["W"], cfg["H"], cfg["den"], cfg["block"]
    kw = dict(use_taa=cfg["taa"], separate_matrices=cfg["sep"], raw_f16=cfg["f16"], position_type=cfg["pos"])
    a, b = oracle.OracleChain(W, H, den, block, **kw), ref.RefChain(W, H, den, block, **kw)
    rng = np.random.default_rng(cfg["seed"])

    def canon(x):
        x = _canon_nan(x)
        if x.dtype == np.float32:
            x = x.copy()
            x[np.isnan(x)] = np.float32(np.nan)
            return x.view(np.uint32)
        return x

    for k, f in enumerate(range(cfg["first"], cfg["first"] + cfg["frames"])):
        with np.errstate(over="ignore"):
            fr = _perturb(synth.render_frame(W, H, f), rng, cfg["modes"][k])
            if den.endswith("x3"):
                sq = second_moment_plane(oracle, a)
                a.average_squared[...] = sq
                b.average_squared[...] = sq
            a.run_frame(f, fr, keep_debug=True)
            b.run_frame(f, fr, keep_debug=True)
        for name in ("motion", "spp", "illum", "prev_depth", "blend_final", "taa_final", "taa_history"):
            np.testing.assert_array_equal(canon(getattr(a, name)), canon(getattr(b, name)), err_msg=f"{name}, frame {f}")
        for blk in a.blocks:
            np.testing.assert_array_equal(canon(a.denoised[blk]), canon(b.denoised[blk]), err_msg=f"denoised b={blk}, frame {f}")
            np.testing.assert_array_equal(a.finals[blk], b.finals[blk], err_msg=f"final b={blk}, frame {f}")
        if den == "bmfr":
            ox, oy = oracle.bmfr_block_offset(block, f)
            Hb, Wb = a.weights.shape[1:]
            inside = np.array([[bx * block - ox < W and bx * block - ox + block > 0 and by * block - oy < H and by * block - oy + block > 0
                                for bx in range(Wb)] for by in range(Hb)])
            np.testing.assert_array_equal(canon(a.weights)[:, inside], canon(b.weights)[:, inside], err_msg=f"weights, frame {f}")
            big = np.repeat(np.repeat(inside, block, 0), block, 1)
            fa, fb = canon(a.features).copy(), canon(b.features).copy()
            for x in (fa, fb):
                x[x == 0x8000] = 0      # a block holding both +0 and -0 depths: which zero a min / max reduction returns depends on its order
            np.testing.assert_array_equal(fa[:, big], fb[:, big], err_msg=f"feature buffer, frame {f}")
