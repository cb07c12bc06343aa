"""Generates the committed golden vectors under tests/golden/.

The reference cannot be executed as it ships (GLSL; no glslc / Vulkan ICD here) and holds no test vectors, so the
goldens are OUTPUTS OF THE REFERENCE'S OWN TEXT compiled for the CPU in this container (oracle/_ref: the shaders through
oracle/glsl_shim, the host conversion code of source/io/RenderIO.cpp through oracle/host_shim) on deterministic inputs:

  <case>.json               per-frame SHA-256 of every plane of the chain.  Produced by the oracle AND by the reference's
                            shader source; the script refuses to write a file on which the two differ, and records which
                            produced it ("verified_against").  Without /root/reference it can only re-derive them from the
                            oracle and says so.
  host_conversions.json     inputs and outputs (bit patterns) of GBufferIO's import / export conversions as computed by
                            the reference's C++ (oracle/_ref/libhostref.so).  Not regenerated without the reference.
  reference_vectors.json    the same for vsg's matrix inverse, Accumulator::set_camera_matrices (the push-constant block of six
                            frames, both matrix modes), formatConverter.comp and the demodulation statements of ptRaygen.rgen.

tests/test_oracle_kat.py compares the oracle with both on every machine, including those where /root/reference and
oracle/_ref do not exist.

    python -m tests.golden.make_golden        # rewrites tests/golden/*.json
"""
import ctypes
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
    # the other block sizes, the single-scale BFR, the WORLD feature modes, the offline (combined-matrix) accumulator with an
    # rgba16f raw input, the BMFR X8X16X32 wiring; frames 7.. so that the negative-jitter frames 8 / 9 are inside
    "bmfr16_taa_208x144_4f": dict(W=208, H=144, denoiser="bmfr", block=16, taa=True, frames=4, first=7),
    "bmfr8_200x136_3f": dict(W=200, H=136, denoiser="bmfr", block=8, taa=False, frames=3, first=7),
    "bfr16_taa_168x104_4f": dict(W=168, H=104, denoiser="bfr", block=16, taa=True, frames=4, first=7),
    "bmfr32_world1_160x128_3f": dict(W=160, H=128, denoiser="bmfr", block=32, taa=False, frames=3, first=7, position_type=1),
    "bmfr32_world2_160x128_3f": dict(W=160, H=128, denoiser="bmfr", block=32, taa=False, frames=3, first=7, position_type=2),
    "bmfr32_combined_f16_192x128_4f": dict(W=192, H=128, denoiser="bmfr", block=32, taa=True, frames=4, separate=False, raw_f16=True),
    "bmfrx3_taa_160x128_3f": dict(W=160, H=128, denoiser="bmfrx3", block=32, taa=True, frames=3),
}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_case(oracle, c):
    orc = oracle.OracleChain(c["W"], c["H"], c["denoiser"], c["block"], use_taa=c["taa"], separate_matrices=c.get("separate", True),
                             raw_f16=c.get("raw_f16", False), position_type=c.get("position_type", 0))
    frames = []
    for f in range(c.get("first", 0), c.get("first", 0) + c["frames"]):
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


def _hex(a):
    a = np.ascontiguousarray(a)
    return {"dtype": str(a.dtype), "shape": list(a.shape), "hex": a.tobytes().hex()}


def from_hex(d):
    return np.frombuffer(bytes.fromhex(d["hex"]), dtype=d["dtype"]).reshape(d["shape"]).copy()


def host_conversion_inputs():
    """small planes with the special values of each conversion (poles, the miss normal, truncation edges, the miss distance)"""
    rng = np.random.default_rng(2026)
    H, W = 6, 9
    cam = synth.render_frame(32, 32, 3).camera
    m64 = np.concatenate([np.asarray(x, np.float32).reshape(-1) for x in (cam.view, cam.inv_view, cam.proj, cam.inv_proj)])
    position = rng.uniform(-30, 30, (H, W, 4)).astype(np.float32)
    v = rng.standard_normal((H, W, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    cart = np.concatenate([v, np.ones((H, W, 1))], -1).astype(np.float32)
    cart.reshape(-1, 4)[:6] = np.float32([[0, 0, 1, 1], [0, 0, -1, 1], [1, 0, 0, 1], [0, -1, 0, 1], [-1, 0, 0, 1], [0.6, 0.8, 0, 1]])
    albedo = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
    albedo.reshape(-1)[:8] = np.float32([0, 0.5, 1.0, 0.999, 0.0039, 0.25, 1 / 255, 254.5 / 255])
    depth = rng.uniform(0.05, 200, (H, W)).astype(np.float32)
    depth.reshape(-1)[:2] = np.float32([0.0, 1e10])
    sph = np.stack([rng.uniform(0, np.pi, (H, W)), rng.uniform(-np.pi, np.pi, (H, W))], -1).astype(np.float32)
    sph.reshape(-1, 2)[:4] = np.float32([[0, 0], [np.pi, 0], [np.pi / 2, np.pi], [0, np.pi / 4]])
    unorm = rng.integers(0, 256, (H, W, 4), dtype=np.uint8)
    unorm.reshape(-1)[:4] = np.uint8([0, 1, 254, 255])
    return dict(matrices64=m64, position=position, cartesian=cart, albedo=albedo, depth=depth, spherical=sph, unorm=unorm)


def host_conversions(R):
    i = host_conversion_inputs()
    d, n, a = R.gbuffer_import(i["matrices64"][16:32], i["position"], i["cartesian"], i["albedo"])
    p, c, u = R.gbuffer_export(i["matrices64"], True, i["depth"], i["spherical"], i["unorm"])
    return {"produced_by": "GBufferIO's conversions as the reference's own C++ text (source/io/RenderIO.cpp:101-120, :160-211, :312-382) "
                           "compiled by oracle/host_shim against vsg's maths headers",
            "inputs": {k: _hex(x) for k, x in i.items()},
            "outputs": {"import_depth": _hex(d), "import_normal": _hex(n), "import_albedo": _hex(a),
                        "export_position": _hex(p), "export_normal": _hex(c), "export_unorm": _hex(u)}}


def more_inputs():
    """inputs of the other reference-pinned functions: matrices for vsg's inverse (affine -> t_inverse_4x3, general,
    singular -> NaN diagonal), an image with the format converter's rounding boundaries, radiance / albedo / position.x
    planes with the demodulation's clamps"""
    rng = np.random.default_rng(77)
    mats = []
    for f in (0, 5):
        cam = synth.camera(160, 128, f)
        mats += [cam.view, cam.proj, cam.inv_view, cam.inv_proj]
    for k in range(16):
        m = rng.uniform(-3, 3, 16).astype(np.float32)
        if k % 2 == 0:
            m[3] = m[7] = m[11] = 0
            m[15] = 1
        if k % 7 == 6:
            m[4:8] = m[0:4]
        mats.append(m)
    mats += [np.zeros(16, np.float32), np.eye(4, dtype=np.float32).reshape(-1)]
    H, W = 5, 12
    img = rng.uniform(-0.25, 1.25, (H, W, 4)).astype(np.float32)
    img[0, 0] = [np.nan, np.inf, -np.inf, 0.5]
    img[0, 1] = [0.0, 1.0, 0.5 / 255.0, 254.5 / 255.0]
    img[1] = (np.arange(W)[:, None] / 255.0 + np.array([0, 1e-4, -1e-4, 0.5 / 255.0])[None, :]).astype(np.float32)
    L = rng.uniform(-1, 14, (H, W, 4)).astype(np.float32)
    L.reshape(-1)[:10] = np.float32([0.0, -0.0, 1e-40, 1e-6, 9.999999, 10.0, 10.000001, 3e38, np.inf, -np.inf])
    alb = rng.uniform(0, 1, (H, W, 4)).astype(np.float32)
    alb[2, :4, :3] = 0.0
    alb.reshape(-1)[40:46] = np.float32([0.0, 1e-7, 1e-6, 1e-3, 0.01, 1.0])
    px = rng.uniform(-50, 50, (H, W)).astype(np.float32)
    px[3, 2:6] = np.inf
    px[4, 3] = -np.inf
    return dict(matrices=np.stack([np.ascontiguousarray(m, np.float32) for m in mats]), image=img,
                image_f16=img.astype(np.float16).view(np.uint16), image_u8=(np.clip(np.nan_to_num(img), 0, 1) * 255).astype(np.uint8),
                radiance=L, albedo=alb, position_x=px)


def push_constant_blocks(O, set_camera_matrices=None):
    """Accumulator::set_camera_matrices over six frames of the synthetic camera, both matrix modes: the oracle's 212-byte
    block (and, given the reference's function, the reference's) per frame"""
    from vulkanpbrt_b200.pipeline import _combined
    W, H = 160, 128
    out = {}
    for separate in (True, False):
        chain = O.OracleChain(W, H, "bmfr", 32, separate_matrices=separate)
        pc = np.zeros(53, np.float32)
        blocks, prev_cam = [], None

        def pack(cam):
            if separate:
                return np.concatenate([cam.view, cam.inv_view, cam.proj, cam.inv_proj]).astype(np.float32), 1
            vp, ivp = _combined(cam)
            return np.concatenate([vp, ivp, np.zeros(32, np.float32)]).astype(np.float32), 0

        for f in range(6):
            cam = synth.camera(W, H, f)
            chain._set_camera_matrices(f, cam)
            got = np.concatenate([np.array(list(chain.pc.view), np.float32), np.array(list(chain.pc.inv_view), np.float32),
                                  np.array(list(chain.pc.prev_view), np.float32), np.array(list(chain.pc.prev_origin), np.float32),
                                  np.array([chain.pc.frame_number], np.int32).view(np.float32)])
            if set_camera_matrices is not None:
                cur64, has = pack(cam)
                prev64 = (np.concatenate([chain.prev_view, np.zeros(48, np.float32)]).astype(np.float32) if separate
                          else pack(prev_cam if prev_cam is not None else cam)[0])
                p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
                assert set_camera_matrices(1 if separate else 0, f, p(cur64), has, p(prev64), has if not separate else 0, p(pc)) == 0
                got = pc.copy()
            blocks.append(got)
            chain.prev_view = np.asarray(cam.view, np.float32).copy()
            chain.prev_cam = prev_cam = cam
        out["separate" if separate else "combined"] = np.stack(blocks)
    return out


def reference_vectors(O, R):
    import ctypes as C
    i = more_inputs()
    h = C.CDLL(str(R._HOST_LIB))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    inv = np.zeros_like(i["matrices"])
    for k, m in enumerate(i["matrices"]):
        h.hostref_inverse(p(np.ascontiguousarray(m)), p(inv[k]))
    pcs = push_constant_blocks(O, h.hostref_set_camera_matrices)
    return {"produced_by": "the reference's own text compiled for the CPU (oracle/_ref): vsg::inverse(mat4) (maths_transform.cpp:36-156) and "
                           "Accumulator::set_camera_matrices (Accumulator.cpp:85-117) through oracle/host_shim; formatConverter.comp and the "
                           "demodulation statements of ptRaygen.rgen:81-88 through oracle/glsl_shim",
            "inputs": {k: _hex(x) for k, x in i.items()},
            "outputs": {"inverse": _hex(inv), "format_converter_f32": _hex(R.format_converter(i["image"])),
                        "format_converter_f16": _hex(R.format_converter(i["image_f16"])), "format_converter_u8": _hex(R.format_converter(i["image_u8"])),
                        "demodulate": _hex(R.demodulate(i["radiance"], i["albedo"], i["position_x"])),
                        "push_constants_separate": _hex(pcs["separate"]), "push_constants_combined": _hex(pcs["combined"])}}


class _RefAsOracle:
    """the reference's shader source behind the oracle's chain interface"""
    def __init__(self, O, R):
        self.OracleChain, self.f16_bits_to_f32 = R.RefChain, O.f16_bits_to_f32


if __name__ == "__main__":
    from oracle import oracle as O
    from oracle import ref as R
    out = Path(__file__).resolve().parent
    have_ref = R.build() and R.build_host()
    for name, c in CASES.items():
        got = run_case(O, c)
        if have_ref:
            ref = run_case(_RefAsOracle(O, R), c)
            if ref["frames"] != got["frames"]:
                raise SystemExit(f"{name}: the oracle differs from the reference's shader source -- nothing written")
            got["verified_against"] = "the reference's shader source (shaders/*.comp via oracle/glsl_shim -> oracle/_ref/libref.so): identical hashes"
        else:
            got["verified_against"] = "nothing (no /root/reference here): derived from the oracle alone"
        (out / f"{name}.json").write_text(json.dumps(got, indent=1))
        print("wrote", name, "--", got["verified_against"])
    if have_ref:
        (out / "host_conversions.json").write_text(json.dumps(host_conversions(R), indent=1))
        (out / "reference_vectors.json").write_text(json.dumps(reference_vectors(O, R), indent=1))
        print("wrote host_conversions, reference_vectors")
