"""The C++ class layer (include/vkpbrt/vkpbrt.hpp) driven like the reference's main(): examples/cpp_frame_loop.cpp is
compiled with g++ and linked against the product library (GPU) or the test emulator (CPU), run on a synthetic
sequence and compared with the oracle bit for bit."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from vulkanpbrt_b200 import synth

ROOT = Path(__file__).resolve().parents[1]


def _run(tmp_path, oracle, libdir, libname, denoiser, taa, W=160, H=128, frames=3, block=32, extra_flags=()):
    exe = tmp_path / "cpp_frame_loop"
    cmd = ["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), *extra_flags, str(ROOT / "examples" / "cpp_frame_loop.cpp"), "-o", str(exe),
           f"-L{libdir}", f"-l{libname}", f"-Wl,-rpath,{libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # block 0 = X8X16X32: three denoisers + blender; the blender's second-moment input is illumination_images[1], which no
    # shader writes (zeros here): sigma = sqrt(0 - av^2) is NaN in the reference too, and the blend that follows is pinned
    orc = oracle.OracleChain(W, H, denoiser + ("x3" if block == 0 else ""), block or 32, use_taa=taa)
    want = []
    for f in range(frames):
        fr = synth.render_frame(W, H, f)
        base = tmp_path / f"frame_{f}"
        fr.depth.tofile(str(base) + ".depth"); fr.normal.tofile(str(base) + ".normal")
        fr.albedo.tofile(str(base) + ".albedo"); fr.illumination.tofile(str(base) + ".illum")
        np.concatenate([fr.camera.view, fr.camera.inv_view, fr.camera.proj, fr.camera.inv_proj]).astype(np.float32).tofile(str(base) + ".cam")
        orc.run_frame(f, fr)
        want.append(orc.final().copy())
    r = subprocess.run([str(exe), str(tmp_path), str(W), str(H), str(frames), denoiser, "1" if taa else "0", str(block)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for f in range(frames):
        got = np.fromfile(tmp_path / f"final_{f}.bgra", dtype=np.uint8).reshape(H, W, 4)
        np.testing.assert_array_equal(got, want[f], err_msg=f"frame {f}")


@pytest.mark.parametrize("denoiser,taa,block", [("bmfr", True, 32), ("bfr", False, 16), ("bmfr", False, 8), ("bfr", True, 32), ("bfr", True, 0)])
def test_the_references_own_wiring_source_compiles_against_the_cpp_layer(tmp_path, oracle, denoiser, taa, block):
    """the drop-in claim, literally: source/util/DenoiserUtils.cpp -- the reference's text, only its #include lines removed --
    is compiled against include/vkpbrt/vkpbrt.hpp behind the three documented substitutions (examples/cpp_frame_loop.cpp
    with -DVKPBRT_REFERENCE_WIRING) and drives the modules; the frames equal the oracle bit for bit -- single block sizes and the BFR
    X8X16X32 case (three denoisers + BFRBlender).  (The BMFR X8X16X32 case of that file indexes illumination_images[2] of a
    two-image buffer and is not exercised.)"""
    import re
    ref_src = Path("/root/reference/source/util/DenoiserUtils.cpp")
    if not ref_src.exists():
        pytest.skip("/root/reference is not mounted")
    wiring = tmp_path / "reference_wiring.inc"             # derived from the reference: lives in the test's temporary directory only
    wiring.write_text(re.sub(r'^\s*#include[^\n]*$', '', ref_src.read_text(), flags=re.M))
    subprocess.run(["make", "-C", str(ROOT / "tests" / "hostsim")], check=True, capture_output=True)
    _run(tmp_path, oracle, ROOT / "tests" / "hostsim", "vkpbrt_hostsim", denoiser, taa, block=block, extra_flags=(f"-DVKPBRT_REFERENCE_WIRING={wiring}",))


@pytest.mark.parametrize("denoiser,taa", [("bmfr", True), ("bfr", False)])
def test_cpp_layer_on_emulator(tmp_path, oracle, denoiser, taa):
    subprocess.run(["make", "-C", str(ROOT / "tests" / "hostsim")], check=True, capture_output=True)
    _run(tmp_path, oracle, ROOT / "tests" / "hostsim", "vkpbrt_hostsim", denoiser, taa)


@pytest.mark.gpu
@pytest.mark.parametrize("denoiser,taa", [("bmfr", True), ("bfr", False)])
def test_cpp_layer_on_gpu(tmp_path, oracle, denoiser, taa):
    _run(tmp_path, oracle, ROOT / "vulkanpbrt_b200" / "lib", "vkpbrt_b200", denoiser, taa, W=640, H=360, frames=4)


@pytest.mark.parametrize("W,H,world,taa", [(1920, 1080, 2, True), (1920, 2160, 2, False), (3840, 2160, 4, True), (3840, 2160, 8, True),
                                           (1920, 8640, 8, False), (7680, 4320, 8, False), (640, 360, 4, True)])
def test_cpp_band_plan_equals_the_python_plan(tmp_path, W, H, world, taa):
    """include/vkpbrt/band_plan.hpp (what a C++ host shards with) against vulkanpbrt_b200/multigpu.py BandPlan: band
    boundaries, owned / accumulated / input rows and every transfer list, over a full jitter period"""
    from vulkanpbrt_b200.multigpu import BandPlan
    exe = tmp_path / "band_plan_dump"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "examples" / "band_plan_dump.cpp"), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    frames = 17
    got = subprocess.run([str(exe), str(W), str(H), str(world), "1" if taa else "0", str(frames)], capture_output=True, text=True).stdout.split("\n")
    try:
        plan = BandPlan(W, H, world, 32, 24, taa)
    except ValueError as e:
        assert got[0].startswith("error"), got[0]
        assert "edge regions" in str(e)
        return
    want = [f"brow {b}" for b in plan.brow]
    want += [f"input {g} {plan.input_rows(g)[0]} {plan.input_rows(g)[1]}" for g in range(world)]
    for f in range(frames):
        for g in range(world):
            o, a = plan.owned_rows(g, f), plan.accumulate_rows(g, f)
            want.append(f"rows {f} {g} {o[0]} {o[1]} {a[0]} {a[1]}")
        want += [f"history {f + 1} {t.src} {t.dst} {t.plane} {t.rows[0]} {t.rows[1]}" for t in plan.history_transfers(f + 1)]
        want += [f"stale {f} {t.src} {t.dst} {t.plane} {t.rows[0]} {t.rows[1]}" for t in plan.stale_column_transfers(f)]
        want += [f"final {f} {t.src} {t.dst} {t.plane} {t.rows[0]} {t.rows[1]}" for t in plan.final_transfers(f)]
    assert [line for line in got if line] == want


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,taa", [(2, True), (3, True), (3, False)])
def test_cpp_banded_ranks_on_emulator(tmp_path, oracle, world, taa):
    """examples/cpp_banded_ranks.cpp: BandPlan + PeerMemory + HaloExchange + the modules' band ranges, ranks as threads
    over the test emulator; the frame assembled from the ranks' owned rows must equal the oracle bit for bit"""
    subprocess.run(["make", "-C", str(ROOT / "tests" / "hostsim")], check=True, capture_output=True)
    libdir = ROOT / "tests" / "hostsim"
    exe = tmp_path / "cpp_banded_ranks"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "examples" / "cpp_banded_ranks.cpp"), "-o", str(exe),
                        f"-L{libdir}", "-lvkpbrt_hostsim", f"-Wl,-rpath,{libdir}", "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    W, H, frames = 64, 416, 19
    orc = oracle.OracleChain(W, H, "bmfr", 32, use_taa=taa)
    want = []
    for f in range(frames):
        fr = synth.render_frame(W, H, f)
        base = tmp_path / f"frame_{f}"
        fr.depth.tofile(str(base) + ".depth"); fr.normal.tofile(str(base) + ".normal")
        fr.albedo.tofile(str(base) + ".albedo"); fr.illumination.tofile(str(base) + ".illum")
        np.concatenate([fr.camera.view, fr.camera.inv_view, fr.camera.proj, fr.camera.inv_proj]).astype(np.float32).tofile(str(base) + ".cam")
        orc.run_frame(f, fr)
        want.append(orc.final().copy())
    r = subprocess.run([str(exe), str(tmp_path), str(W), str(H), str(frames), str(world), "1" if taa else "0"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    for f in range(frames):
        got = np.fromfile(tmp_path / f"final_{f}.bgra", dtype=np.uint8).reshape(H, W, 4)
        np.testing.assert_array_equal(got, want[f], err_msg=f"frame {f}")


def test_cpp_layer_is_clean_under_the_sanitizers(tmp_path, oracle):
    """the header-only layer (closures that keep modules alive, borrowed vs owned handles, the X8X16X32 wiring with its
    side lanes) with -fsanitize=address,undefined and leak checking on the example's own code: no use-after-free, no
    leak, no undefined behaviour -- the lifetime rules of DESIGN.md section 1 hold as implemented"""
    import os
    subprocess.run(["make", "-C", str(ROOT / "tests" / "hostsim")], check=True, capture_output=True)
    libdir = ROOT / "tests" / "hostsim"
    exe = tmp_path / "cpp_frame_loop_asan"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I", str(ROOT / "include"),
                        str(ROOT / "examples" / "cpp_frame_loop.cpp"), "-o", str(exe), f"-L{libdir}", "-lvkpbrt_hostsim", f"-Wl,-rpath,{libdir}"],
                       capture_output=True, text=True)
    if r.returncode != 0 and ("asan" in r.stderr or "ubsan" in r.stderr):
        pytest.skip("this toolchain has no sanitizer runtime")
    assert r.returncode == 0, r.stderr
    W, H, frames = 96, 64, 2
    for f in range(frames):
        fr = synth.render_frame(W, H, f)
        base = tmp_path / f"frame_{f}"
        fr.depth.tofile(str(base) + ".depth"); fr.normal.tofile(str(base) + ".normal")
        fr.albedo.tofile(str(base) + ".albedo"); fr.illumination.tofile(str(base) + ".illum")
        np.concatenate([fr.camera.view, fr.camera.inv_view, fr.camera.proj, fr.camera.inv_proj]).astype(np.float32).tofile(str(base) + ".cam")
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:halt_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    for args in (("bmfr", "1", "32"), ("bfr", "1", "0")):
        r = subprocess.run([str(exe), str(tmp_path), str(W), str(H), str(frames), *args], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0 and "ERROR" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
