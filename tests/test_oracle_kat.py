"""CPU tests of the oracle itself: known answers derived independently from the reference's constants
and formulas (the reference ships no test vectors), plus the committed golden
hashes that pin the oracle's outputs across machines and revisions (tests/golden/make_golden.py)."""
import ctypes as C
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from vulkanpbrt_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden"


def test_f16_roundtrip_all_halves(oracle):
    L = oracle.lib()
    bits = np.arange(65536, dtype=np.uint16)
    ref = bits.view(np.float16).astype(np.float32)
    for b in range(0, 65536, 7):
        got = L.vkpbrt_oracle_f16_to_f32(int(b))
        if np.isnan(ref[b]):
            assert np.isnan(got)
        else:
            assert got == ref[b]
            assert L.vkpbrt_oracle_f32_to_f16(float(ref[b])) == b


def test_f32_to_f16_rounds_to_nearest_even(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.standard_normal(4000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 5, 4000).astype(np.float32),
        np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e-8, 5.96e-8, 2.98e-8, 2.9802322e-8, 6.1e-5, 6.1035156e-5,
                  1.0009765, 1.00048828125, 1.0014648], np.float32)])
    with np.errstate(over="ignore"):
        want = vals.astype(np.float16).view(np.uint16)
    for v, w in zip(vals, want):
        assert L.vkpbrt_oracle_f32_to_f16(float(v)) == int(w), float(v)


def test_unorm8(oracle):
    L = oracle.lib()
    assert L.vkpbrt_oracle_f32_to_unorm8(float("nan")) == 0
    assert L.vkpbrt_oracle_f32_to_unorm8(-3.0) == 0
    assert L.vkpbrt_oracle_f32_to_unorm8(7.0) == 255
    assert L.vkpbrt_oracle_f32_to_unorm8(1.0 / 256.0) == 1       # first-frame sample count (accumulator.comp:43)
    for c in range(256):
        assert L.vkpbrt_oracle_f32_to_unorm8(c / 255.0) == c


def _round_f32(x):
    """Fraction -> nearest-even binary32, exactly"""
    from fractions import Fraction
    if x == 0:
        return np.float32(0.0)
    near = np.float32(float(x))
    best = None
    for c in (near, np.nextafter(near, np.float32(np.inf)), np.nextafter(near, np.float32(-np.inf))):
        d = abs(Fraction(float(c)) - x)
        even = (int(np.float32(c).view(np.uint32)) & 1) == 0
        if best is None or d < best[0] or (d == best[0] and even):
            best = (d, np.float32(c))
    return best[1]


def test_kernel_unorm8_decode_is_the_exact_quotient_for_all_codes():
    """csrc/common.cuh unorm8_scale: fma(c, RH, c * RL) with 1/255 = RH + RL must equal the oracle's c / 255.0f
    (an IEEE division) for every code; the constants are read from the source so the test follows the kernel"""
    import re
    from fractions import Fraction
    src = (Path(__file__).resolve().parents[1] / "vulkanpbrt_b200" / "csrc" / "common.cuh").read_text()
    m = re.search(r"fmaf\(cf, (-?0x[0-9a-fA-F.]+p[-+]?\d+)f, mul_rn\(cf, (-?0x[0-9a-fA-F.]+p[-+]?\d+)f\)\)", src)
    assert m, "unorm8_scale not found in common.cuh"
    rh, rl = (Fraction(float.fromhex(g)) for g in m.groups())
    assert float(np.float32(float(rh))) == float(rh) and float(np.float32(float(rl))) == float(rl)   # binary32 constants
    for c in range(256):
        t = _round_f32(Fraction(c) * rl)
        q = _round_f32(Fraction(c) * rh + Fraction(float(t)))
        assert q == np.float32(c) / np.float32(255.0), c


def test_block_coordinate_division_by_reciprocal_is_exact():
    """bmfr.cu evaluates i / (B - 1), i = 0 .. B-1, as q0 = i*r, e = fma(-q0, B-1, i), q = fma(e, r, q0) with r = RN(1/(B-1)):
    must equal the IEEE quotient the oracle computes, for every block size"""
    from fractions import Fraction
    for B in (8, 16, 32):
        b = Fraction(B - 1)
        r = Fraction(float(np.float32(1.0) / np.float32(B - 1)))
        for i in range(B):
            a = Fraction(i)
            q0 = Fraction(float(_round_f32(a * r)))
            e = Fraction(float(_round_f32(a - q0 * b)))
            q = _round_f32(e * r + q0)
            assert q == np.float32(i) / np.float32(B - 1), (B, i)


# ivec2(vec2(b, b) * pixelOffsets[f % 16]) for the table of bmfrGeneral.comp:36
BMFR_OFFSETS = {
    32: [(22, 27), (30, 16), (13, 24), (31, 0), (11, 18), (0, 11), (25, 14), (0, 24), (11, -2), (-1, 0), (30, 3), (27, 19),
         (1, 3), (13, 5), (0, 16), (23, 12)],
    16: [(11, 13), (15, 8), (6, 12), (15, 0), (5, 9), (0, 5), (12, 7), (0, 12), (5, -1), (0, 0), (15, 1), (13, 9), (0, 1),
         (6, 2), (0, 8), (11, 6)],
    8: [(5, 6), (7, 4), (3, 6), (7, 0), (2, 4), (0, 2), (6, 3), (0, 6), (2, 0), (0, 0), (7, 0), (6, 4), (0, 0), (3, 1),
        (0, 4), (5, 3)],
}
BFR_OFFSETS = [(-7, -11), (-14, -8), (-5, -12), (-15, -1), (-5, -9), (-1, -4), (-14, -7), (0, -13), (-5, -1), (-1, 0),
               (-15, -2), (-14, -10), (-1, -1), (-6, -3), (0, -8), (-10, -4)]


@pytest.mark.parametrize("b", [8, 16, 32])
def test_bmfr_block_jitter_table(oracle, b):
    for f in range(48):
        assert oracle.bmfr_block_offset(b, f) == BMFR_OFFSETS[b][f % 16]


def test_bfr_block_jitter_table(oracle):
    for f in range(40):
        assert oracle.bfr_block_offset(f) == BFR_OFFSETS[f % 16]


def _ref_hash(a):
    """bmfrGeneral.comp:103-113, restated independently with Python integers"""
    M = 0xFFFFFFFF
    a = ((a + 0x7ed55d16) + (a << 12)) & M
    a = ((a ^ 0xc761c23c) ^ (a >> 19)) & M
    a = ((a + 0x165667b1) + (a << 5)) & M
    a = ((a + 0xd3a2646c) ^ (a << 9)) & M
    a = ((a + 0xfd7046c5) + (a << 3)) & M
    a = ((a ^ 0xb55a4f09) ^ (a >> 16)) & M
    return np.float32(a) / np.float32(4294967295.0)      # float(0xffffffff) rounds to 2^32


def test_noise_hash_known_answers(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(7)
    seeds = [0, 1, 255, 1023, 1 << 20, 13 * (1 << 20) * 157 & 0xFFFFFFFF, 0xFFFFFFFF] + [int(s) for s in rng.integers(0, 2**32, 500)]
    for s in seeds:
        assert L.vkpbrt_oracle_bmfr_random(s) == _ref_hash(s)


def test_mat_inverse_matches_numpy(oracle):
    cam = synth.camera(640, 360, 5)
    for m in (cam.view, cam.proj, cam.inv_view):
        inv = oracle.mat_inverse(m).reshape(4, 4).T
        want = np.linalg.inv(np.asarray(m, np.float64).reshape(4, 4).T)
        np.testing.assert_allclose(inv, want, rtol=2e-5, atol=2e-6)


def test_deterministic_transcendentals(oracle):
    L = oracle.lib()
    L.vkpbrt_oracle_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.vkpbrt_oracle_pow.argtypes = [C.c_float, C.c_float]
    L.vkpbrt_oracle_pow.restype = C.c_float
    for x in np.linspace(-7.0, 7.0, 4001).astype(np.float32):
        s, c = C.c_float(), C.c_float()
        L.vkpbrt_oracle_sincos(float(x), C.byref(s), C.byref(c))
        assert abs(s.value - np.sin(np.float64(x))) < 2e-7 and abs(c.value - np.cos(np.float64(x))) < 2e-7
    s, c = C.c_float(), C.c_float()
    L.vkpbrt_oracle_sincos(0.0, C.byref(s), C.byref(c))
    assert (s.value, c.value) == (0.0, 1.0)               # flat ground: theta = 0 -> n = (0, 0, 1) exactly
    for x in np.logspace(-6, 1.1, 2001).astype(np.float32):
        v = L.vkpbrt_oracle_pow(float(x), 0.454545)
        t = np.float64(x) ** np.float64(np.float32(0.454545))
        assert abs(v - t) / t < 2e-6
    assert L.vkpbrt_oracle_pow(0.0, 0.454545) == 0.0 and L.vkpbrt_oracle_pow(1.0, 0.454545) == 1.0


def test_fit_solves_the_least_squares_problem(oracle):
    """bmfrFit's Householder QR + back substitution must reproduce a float64 least-squares fit of the same
    (noisy, fp16) system on the block's own features (checks the algorithm, not the rounding)."""
    W = H = 128
    orc = oracle.OracleChain(W, H, "bmfr", 32)
    fr = synth.render_frame(W, H, 0)
    orc.run_frame(0, fr, keep_debug=True)
    L = oracle.lib()
    feat, wts = orc.features, orc.weights
    checked = 0
    for by in range(1, 4):
        for bx in range(1, 4):
            i = np.arange(1024)
            px, py = i // 32 + bx * 32, i % 32 + by * 32          # bmfrFit.comp:18-19: x = i / 32
            A = np.zeros((1024, 13))
            for c in range(13):
                v = oracle.f16_bits_to_f32(feat[c, py, px])
                if c < 10:
                    r = np.array([L.vkpbrt_oracle_bmfr_random(int(s)) for s in (i + c * (1 << 20))], np.float32)
                    v = v + np.float32(2e-4) * (r - np.float32(0.5))
                A[:, c] = v
            if np.abs(A[:, 10:]).max() == 0:
                continue
            sol = np.linalg.lstsq(A[:, :10], A[:, 10:], rcond=None)[0]
            w = wts[:, by, bx].reshape(10, 3).astype(np.float64)
            # compare the FITS (predictions on the fitted system): the weights themselves are ill-conditioned
            np.testing.assert_allclose(A[:, :10] @ w, A[:, :10] @ sol, atol=2e-3)
            checked += 1
    assert checked >= 4


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _golden_cases():
    from tests.golden.make_golden import CASES
    return sorted(CASES)


@pytest.mark.parametrize("name", _golden_cases())
def test_oracle_matches_committed_golden_hashes(oracle, name):
    """tests/golden/*.json were produced by tests/golden/make_golden.py from this oracle; they pin its
    outputs bit for bit (any change to the restatement, the compiler flags or the host libm shows up here)."""
    from tests.golden.make_golden import CASES, run_case
    spec = json.loads((GOLDEN / f"{name}.json").read_text())
    got = run_case(oracle, CASES[name])
    assert got["frames"] == spec["frames"]


def test_committed_goldens_are_outputs_of_the_reference():
    """the hash files say what produced them: make_golden.py writes them only when the reference's shader source yields the
    same hashes as the oracle"""
    for name in _golden_cases():
        assert "reference's shader source" in json.loads((GOLDEN / f"{name}.json").read_text())["verified_against"]


def test_oracle_matches_the_reference_host_code_vectors(oracle):
    """tests/golden/host_conversions.json holds inputs and outputs of GBufferIO's import / export conversions computed by
    the reference's own C++ (source/io/RenderIO.cpp through oracle/host_shim, written by make_golden.py where the
    reference is mounted): the oracle reproduces every output bit for bit -- on machines without /root/reference too"""
    from tests.golden.make_golden import from_hex, host_conversion_inputs
    spec = json.loads((GOLDEN / "host_conversions.json").read_text())
    i = {k: from_hex(v) for k, v in spec["inputs"].items()}
    for k, v in host_conversion_inputs().items():                      # the committed inputs are the generator's
        np.testing.assert_array_equal(i[k].view(np.uint8), np.ascontiguousarray(v).view(np.uint8), err_msg=k)
    want = {k: from_hex(v) for k, v in spec["outputs"].items()}
    d, n, a = oracle.gbuffer_import(i["matrices64"][16:32], i["position"], i["cartesian"], i["albedo"])
    p, c, u = oracle.gbuffer_export(i["matrices64"][16:32], i["matrices64"][48:64], i["depth"], i["spherical"], i["unorm"])
    for name, got in (("import_depth", d), ("import_normal", n), ("import_albedo", a), ("export_position", p), ("export_normal", c), ("export_unorm", u)):
        np.testing.assert_array_equal(np.ascontiguousarray(got).view(np.uint8), want[name].view(np.uint8), err_msg=name)


def test_oracle_matches_the_reference_vectors(oracle):
    """tests/golden/reference_vectors.json: vsg's matrix inverse, Accumulator::set_camera_matrices (push-constant blocks of
    six frames, both matrix modes), formatConverter.comp and the demodulation statements of ptRaygen.rgen as computed by
    the reference's own text (oracle/_ref, written by make_golden.py) -- the oracle reproduces every output bit for bit
    (NaNs compared as NaNs), with no reference tree on the machine"""
    from tests.golden.make_golden import from_hex, more_inputs, push_constant_blocks
    spec = json.loads((GOLDEN / "reference_vectors.json").read_text())
    i = {k: from_hex(v) for k, v in spec["inputs"].items()}
    for k, v in more_inputs().items():
        np.testing.assert_array_equal(i[k].view(np.uint8), np.ascontiguousarray(v).view(np.uint8), err_msg=k)
    want = {k: from_hex(v) for k, v in spec["outputs"].items()}

    def same(got, name):
        got, w = np.ascontiguousarray(got), want[name]
        assert got.shape == w.shape and got.dtype == w.dtype, name
        if got.dtype == np.float32:
            nan = np.isnan(got) & np.isnan(w)
            np.testing.assert_array_equal(np.where(nan, 0, got.view(np.uint32)), np.where(nan, 0, w.view(np.uint32)), err_msg=name)
        else:
            np.testing.assert_array_equal(got, w, err_msg=name)

    same(np.stack([np.asarray(oracle.vsg_inverse(m), np.float32) for m in i["matrices"]]), "inverse")
    same(oracle.format_converter(i["image"]), "format_converter_f32")
    same(oracle.format_converter(i["image_f16"]), "format_converter_f16")
    same(oracle.format_converter(i["image_u8"]), "format_converter_u8")
    same(oracle.demodulate(i["radiance"], i["albedo"], i["position_x"]), "demodulate")
    pcs = push_constant_blocks(oracle)
    same(pcs["separate"], "push_constants_separate")
    same(pcs["combined"], "push_constants_combined")


# ---- the transposing shuffle networks of bmfr.cu / bfr.cu, restated symbolically ----------------------
def _network(n_values, lanes=32, plain_tail=False):
    """ReduceN<N, 16> of bfr.cu, or with plain_tail MultiReduce<N, 16> of bmfr.cu (N a power of two; once one value
    is left every lane keeps adding its partner's copy).  Returns per lane the expression tree it ends with; a sum is
    a frozenset of its two operands (IEEE addition is commutative, not associative)."""
    vals = [[("x", lane, j) for j in range(n_values)] for lane in range(lanes)]
    n, off = n_values, lanes // 2
    while off >= 1:
        h = (n + 1) // 2
        new = []
        for lane in range(lanes):
            up = bool(lane & off)
            row = []
            if plain_tail and n == 1:
                new.append([frozenset([vals[lane][0], vals[lane ^ off][0]])])
                continue
            for j in range(h):
                mine, theirs = vals[lane], vals[lane ^ off]
                zero = ("zero",)
                keep = (mine[j + h] if j + h < n else zero) if up else mine[j]
                recv = (theirs[j + h] if j + h < n else zero) if up else theirs[j]     # the partner sends what I keep
                row.append(frozenset([keep, recv]) if keep != recv else ("dbl", keep))
            new.append(row)
        vals, n, off = new, h, off // 2
    return [v[0] for v in vals]


def _butterfly(j, lanes=32):
    x = [("x", lane, j) for lane in range(lanes)]
    off = lanes // 2
    while off >= 1:
        x = [frozenset([x[lane], x[lane ^ off]]) for lane in range(lanes)]
        off //= 2
    return x


def _slot22(lane):      # bfr.cu reduce22_slot
    p3 = (lane & 1) + 2 * ((lane >> 1) & 1)
    p1 = p3 + 3 * ((lane >> 2) & 1) + 6 * ((lane >> 3) & 1)
    return p1 + 11 * ((lane >> 4) & 1) if (p3 < 3 and p1 < 11) else -1


def test_transposing_networks_sum_along_the_butterfly_tree():
    """every total leaves the network as exactly the expression tree of a plain xor-butterfly subgroupAdd, and the
    lane -> slot maps the kernels use are the right ones"""
    got = _network(22)
    slots = [_slot22(lane) for lane in range(32)]
    assert sorted(s for s in slots if s >= 0) == list(range(22))
    for lane, s in enumerate(slots):
        if s >= 0:
            assert got[lane] == _butterfly(s)[lane], (lane, s)
    for n in (16, 8, 4, 2):             # bmfr.cu MultiReduce<KP, 16>: value j on the lanes with lane >> (5 - log2 KP) == j
        got = _network(n, plain_tail=True)
        shift = 5 - n.bit_length() + 1
        for lane in range(32):
            j = lane >> shift
            assert got[lane] == _butterfly(j)[lane], (n, lane)
