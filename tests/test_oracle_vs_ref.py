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
