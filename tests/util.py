"""helpers shared by the parity tests"""
import numpy as np

from vulkanpbrt_b200 import DenoisePipeline, DenoisingBlockSize, DenoisingType, synth

BS = {8: DenoisingBlockSize.X8, 16: DenoisingBlockSize.X16, 32: DenoisingBlockSize.X32}


def make_pair(oracle, W, H, denoiser="bmfr", block=32, use_taa=False, separate_matrices=True, raw_f16=False,
              fix_taa_swizzle=False, debug=False, position_type=0):
    """(CUDA pipeline, oracle chain) configured identically"""
    x3 = denoiser.endswith("x3")
    dt = DenoisingType.BMFR if denoiser.startswith("bmfr") else DenoisingType.BFR
    bs = DenoisingBlockSize.X8X16X32 if x3 else BS[block]
    pipe = DenoisePipeline(W, H, dt, bs, use_taa=use_taa, separate_matrices=separate_matrices, raw_f16=raw_f16,
                           fix_taa_swizzle=fix_taa_swizzle, bmfr_debug_outputs=debug, average_squared=x3, position_type=position_type)
    orc = oracle.OracleChain(W, H, denoiser, block, use_taa=use_taa, separate_matrices=separate_matrices, raw_f16=raw_f16,
                             fix_taa_swizzle=fix_taa_swizzle, position_type=position_type)
    return pipe, orc


def second_moment_plane(oracle, orc):
    """a defined averageSquared plane for the blender: E[x^2] >= E[x]^2"""
    av = oracle.f16_bits_to_f32(orc.prev_illu)
    return np.ascontiguousarray((av * av * 1.5 + 0.01).astype(np.float16).view(np.uint16))


def step_both(oracle, pipe, orc, W, H, f, keep_debug=False):
    fr = synth.render_frame(W, H, f)
    if pipe.average_squared_image is not None:
        sq = second_moment_plane(oracle, orc)
        orc.average_squared[...] = sq
        pipe.average_squared_image.upload(sq)
    pipe.run_frame(f, fr)
    pipe.ctx.synchronize()
    orc.run_frame(f, fr, keep_debug=keep_debug)
    return fr


def assert_frame_equal(pipe, orc, f):
    """every plane the frame produced, bit for bit.  After copy_to_back_images the CUDA side's "current"
    accumulate outputs live in the prev_* handles (pointer swap)."""
    acc = pipe.accumulation_buffer
    np.testing.assert_array_equal(acc.motion.download(), orc.motion, err_msg=f"motion, frame {f}")
    np.testing.assert_array_equal(acc.prev_spp.download(), orc.spp, err_msg=f"spp, frame {f}")
    np.testing.assert_array_equal(acc.prev_illu.download(), orc.illum, err_msg=f"accumulated illumination, frame {f}")
    np.testing.assert_array_equal(acc.prev_depth.download(), orc.prev_depth, err_msg=f"prev_depth, frame {f}")
    for m, b in zip([m for m in pipe.modules if hasattr(m, "denoised")], orc.blocks):
        np.testing.assert_array_equal(m.denoised.download(), orc.denoised[b], err_msg=f"denoised history b={b}, frame {f}")
        np.testing.assert_array_equal(m.get_final_descriptor_image().download(), orc.finals[b], err_msg=f"final b={b}, frame {f}")
    np.testing.assert_array_equal(pipe.denoiser_final.download(), orc.denoiser_final(), err_msg=f"denoiser final, frame {f}")
    np.testing.assert_array_equal(pipe.final.download(), orc.final(), err_msg=f"final image, frame {f}")


def tolerance_report(oracle, got_f16_bits, want_f16_bits):
    """north-star radiance tolerance: per-pixel relative error <= 1e-3 on >= 99.9 % of the pixels and
    PSNR >= 60 dB (BASELINE.json).  Returns (fraction within 1e-3, psnr)."""
    a = oracle.f16_bits_to_f32(got_f16_bits)[..., :3].astype(np.float64)
    b = oracle.f16_bits_to_f32(want_f16_bits)[..., :3].astype(np.float64)
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
    frac = float((rel.max(axis=-1) <= 1e-3).mean())
    return frac, psnr(a, b, max(float(b.max()), 1e-6))


def psnr(a, b, peak):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)
