"""tools/vk_oracle (SURVEY.md section 8(c) "second oracle": a raw-Vulkan host for the reference's compiled shaders,
meant for lavapipe) executed against tests/vkmock, whose vkCmdDispatch runs the reference's shader SOURCE
(oracle/_ref/libref.so, compiled by oracle/glsl_shim) behind the Vulkan API.  Its planes must equal the oracle bit for
bit: that checks the host's binding numbers, descriptor types, formats, sampler, specialisation constants,
push-constant blocks, dispatch sizes and copies against the reference's host code.  No real Vulkan driver exists on
either machine of this project; see tools/vk_oracle/README.md for what that leaves untested."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests.test_vk_interop import _vulkan_include
from tests.util import second_moment_plane
from vulkanpbrt_b200 import synth

ROOT = Path(__file__).resolve().parents[1]
SHADERS = ["accumulator_sep", "bmfrPre", "bmfrFit", "bmfrPost", "bfr", "bfrBlender", "taa"]


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    inc = _vulkan_include()
    if inc is None:
        pytest.skip("no vulkan_core.h on this machine")
    from oracle import ref as R
    if not R.build():
        pytest.skip("oracle/_ref is not built and /root/reference is not mounted")
    out = tmp_path_factory.mktemp("vkoracle")
    mock, exe = out / "libvkmock.so", out / "vk_oracle"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-fPIC", "-shared", "-fvisibility=hidden", "-I", str(inc),
                        str(ROOT / "tests" / "vkmock" / "vkmock.cpp"), "-o", str(mock), "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-ffp-contract=off", "-I", str(inc), str(ROOT / "tools" / "vk_oracle" / "vk_oracle.cpp"),
                        "-o", str(exe), "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    spv = out / "spv"
    spv.mkdir()
    for name in SHADERS:                      # stand-ins for SPIR-V: the mock looks the shader up in libref.so by name
        blob = f"VKMOCK-SHADER:{name}\n".encode()
        (spv / f"{name}.comp.spv").write_bytes(blob + b"\0" * (-len(blob) % 4))
    return mock, exe, spv


@pytest.mark.parametrize("W,H,first,frames,den,block,taa", [(160, 128, 0, 3, "bmfr", 32, True), (168, 104, 7, 3, "bfr", 16, False),
                                                            (136, 72, 7, 3, "bmfr", 8, True), (96, 72, 0, 3, "bfrx3", 32, True),
                                                            (96, 72, 7, 2, "bmfrx3", 32, False)])
def test_raw_vulkan_host_replays_the_reference_frame(tmp_path, oracle, built, W, H, first, frames, den, block, taa):
    mock, exe, spv = built
    env = dict(os.environ, VK_ORACLE_LOADER=str(mock), VKMOCK_LIBREF=str(ROOT / "oracle" / "_ref" / "libref.so"))
    x3 = den.endswith("x3")
    orc = oracle.OracleChain(W, H, den, block, use_taa=taa)
    frames_data = []
    for f in range(first, first + frames):
        fr = synth.render_frame(W, H, f)
        base = tmp_path / f"frame_{f}"
        fr.depth.tofile(str(base) + ".depth"); fr.normal.tofile(str(base) + ".normal")
        fr.albedo.tofile(str(base) + ".albedo"); fr.illumination.tofile(str(base) + ".illum")
        np.concatenate([fr.camera.view, fr.camera.inv_view, fr.camera.proj, fr.camera.inv_proj]).astype(np.float32).tofile(str(base) + ".cam")
        if x3:
            # a defined averageSquared plane for the blender, derived from the history the oracle holds
            # BEFORE this frame -- so the oracle has to advance frame by frame while the files are written
            sq = second_moment_plane(oracle, orc)
            sq.tofile(str(base) + ".avgsq")
            orc.average_squared[...] = sq
        orc.run_frame(f, fr)
        frames_data.append({k: np.copy(v) for k, v in dict(final=orc.final(), motion=orc.motion, spp=orc.spp, illum=orc.illum).items()}
                           | {f"denoised{b}": orc.denoised[b].copy() for b in orc.blocks})
    r = subprocess.run([str(exe), str(spv), str(tmp_path), str(tmp_path), str(W), str(H), str(first), str(frames), den, str(block), "1" if taa else "0"],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    for i, f in enumerate(range(first, first + frames)):
        want = frames_data[i]
        got = dict(final=np.fromfile(tmp_path / f"final_{f}.bgra", np.uint8).reshape(H, W, 4),
                   motion=np.fromfile(tmp_path / f"motion_{f}.rg16f", np.uint16).reshape(H, W, 2),
                   spp=np.fromfile(tmp_path / f"spp_{f}.r8", np.uint8).reshape(H, W),
                   illum=np.fromfile(tmp_path / f"illum_{f}.rgba16f", np.uint16).reshape(H, W, 4))
        for b in orc.blocks:
            got[f"denoised{b}"] = np.fromfile(tmp_path / f"denoised{b}_{f}.rgba16f", np.uint16).reshape(2, H, W, 4)
        for k in got:
            np.testing.assert_array_equal(got[k], want[k], err_msg=f"{k}, frame {f}")


def test_real_spirv_is_refused_by_the_mock(tmp_path, built):
    """the mock cannot execute SPIR-V: a run against it with real shader binaries must fail loudly, not pass vacuously"""
    mock, exe, spv = built
    real = tmp_path / "spv"
    real.mkdir()
    for name in SHADERS:
        (real / f"{name}.comp.spv").write_bytes(np.array([0x07230203, 0x00010400, 0, 1, 0], np.uint32).tobytes())
    env = dict(os.environ, VK_ORACLE_LOADER=str(mock), VKMOCK_LIBREF=str(ROOT / "oracle" / "_ref" / "libref.so"))
    r = subprocess.run([str(exe), str(real), str(tmp_path), str(tmp_path), "64", "64", "0", "1", "bmfr", "32", "0"], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode != 0
    assert "real SPIR-V" in r.stderr
