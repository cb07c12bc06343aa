"""Vulkan interop (SURVEY.md section 8(b) "Sync with Vulkan", 8(f)-2): include/vkpbrt/vk_interop.hpp and the C ABI's
vkpbrt_import_external_{memory,semaphore}_fd, executed.

Neither machine of this project has a Vulkan loader, driver or lavapipe, so the Vulkan side is tests/vkmock (exported
memory and semaphores are memfds, see its header) and the CUDA side is the test emulator, whose
cudaImportExternalMemory / cudaImportExternalSemaphore map those memfds.  What runs is the real host code: the interop
header, examples/cpp_vulkan_interop.cpp (a raw-Vulkan renderer stand-in wired to the reference's module sequence) and
the import entry points of api.cpp.  The Vulkan headers are Khronos' vulkan_core.h as vendored in the reference tree
(external/vsgXchange/src/ktx/libktx/dfdutils/vulkan) or a system copy; without either the tests skip.
"""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from vulkanpbrt_b200 import synth

ROOT = Path(__file__).resolve().parents[1]
HOSTSIM = ROOT / "tests" / "hostsim"


def _vulkan_include():
    for cand in (Path("/root/reference/external/vsgXchange/src/ktx/libktx/dfdutils"), Path("/usr/include"), Path("/usr/local/include")):
        if (cand / "vulkan" / "vulkan_core.h").exists():
            return cand
    return None


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    inc = _vulkan_include()
    if inc is None:
        pytest.skip("no vulkan_core.h on this machine")
    out = tmp_path_factory.mktemp("vkinterop")
    subprocess.run(["make", "-C", str(HOSTSIM)], check=True, capture_output=True)
    mock = out / "libvkmock.so"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-fPIC", "-shared", "-fvisibility=hidden", "-I", str(inc),
                        str(ROOT / "tests" / "vkmock" / "vkmock.cpp"), "-o", str(mock), "-ldl"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exes = {}
    for name, src in (("cpp_vulkan_interop", ROOT / "examples" / "cpp_vulkan_interop.cpp"), ("interop_checks", ROOT / "tests" / "vkmock" / "interop_checks.cpp")):
        exe = out / name
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-I", str(inc), str(src), "-o", str(exe),
                            f"-L{HOSTSIM}", "-lvkpbrt_hostsim", f"-Wl,-rpath,{HOSTSIM}", "-ldl", "-lpthread"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        exes[name] = exe
    return mock, exes


def _env(mock, **extra):
    return dict(os.environ, VKPBRT_VULKAN_LIBRARY=str(mock), **extra)


def _write_sequence(tmp_path, oracle, W, H, frames, denoiser, taa):
    orc = oracle.OracleChain(W, H, denoiser, 32, use_taa=taa)
    want = []
    for f in range(frames):
        fr = synth.render_frame(W, H, f)
        base = tmp_path / f"frame_{f}"
        fr.depth.tofile(str(base) + ".depth"); fr.normal.tofile(str(base) + ".normal")
        fr.albedo.tofile(str(base) + ".albedo"); fr.illumination.tofile(str(base) + ".illum")
        np.concatenate([fr.camera.view, fr.camera.inv_view, fr.camera.proj, fr.camera.inv_proj]).astype(np.float32).tofile(str(base) + ".cam")
        orc.run_frame(f, fr)
        want.append(orc.final().copy())
    return want


@pytest.mark.parametrize("denoiser,taa,dedicated", [("bmfr", True, "0"), ("bfr", False, "1")])
def test_frames_shared_with_a_vulkan_renderer_equal_the_oracle(tmp_path, oracle, built, denoiser, taa, dedicated):
    """G-buffer and illumination travel Vulkan image -> exported plane -> kernels, the denoised frame travels back
    plane -> Vulkan image -> host, ordered only by the two exported timeline semaphores; every frame must equal the
    oracle bit for bit.  dedicated=1: the mock exports buffers only from dedicated allocations."""
    mock, exes = built
    W, H, frames = 160, 128, 4
    want = _write_sequence(tmp_path, oracle, W, H, frames, denoiser, taa)
    r = subprocess.run([str(exes["cpp_vulkan_interop"]), str(tmp_path), str(W), str(H), str(frames), denoiser, "1" if taa else "0"],
                       capture_output=True, text=True, env=_env(mock, VKMOCK_DEDICATED_ONLY=dedicated), timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for f in range(frames):
        got = np.fromfile(tmp_path / f"final_{f}.bgra", dtype=np.uint8).reshape(H, W, 4)
        np.testing.assert_array_equal(got, want[f], err_msg=f"frame {f}")


def test_a_vulkan_device_that_is_not_the_cuda_device_is_refused(tmp_path, oracle, built):
    """external memory can only be imported on the exporting GPU: the host picks the CUDA device by UUID and fails
    loudly when none matches"""
    mock, exes = built
    _write_sequence(tmp_path, oracle, 64, 64, 1, "bmfr", False)
    r = subprocess.run([str(exes["cpp_vulkan_interop"]), str(tmp_path), "64", "64", "1", "bmfr", "0"], capture_output=True, text=True,
                       env=_env(mock, VKMOCK_DEVICE_UUID="SOME-OTHER-GPU-0"), timeout=120)
    assert r.returncode == 1
    assert "no CUDA device has the UUID" in r.stderr


@pytest.mark.parametrize("dedicated", ["0", "1"])
def test_interop_edge_cases(built, dedicated):
    """tests/vkmock/interop_checks.cpp: missing extensions / features reported by name, import argument checks, a plane
    is the same bytes on both sides (both directions, through TILING_OPTIMAL images), no file descriptor leaks, a
    timeline semaphore blocks and releases a waiter across the two APIs, the frame object's handshake values"""
    mock, exes = built
    r = subprocess.run([str(exes["interop_checks"])], capture_output=True, text=True, env=_env(mock, VKMOCK_DEDICATED_ONLY=dedicated), timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert [l for l in r.stdout.split("\n") if l] == ["ok api_load_reports_missing_entry_points", "ok device_uuid_match", "ok import_argument_checks",
                                                      "ok shared_plane_round_trip_no_fd_leak", "ok image_copy_record", "ok shared_timeline_both_sides",
                                                      "ok shared_frame"]


def test_interop_header_is_inert_without_vulkan_headers(tmp_path):
    """compile guard: with no Vulkan headers on the include path the header defines VKPBRT_HAVE_VULKAN 0 and nothing else"""
    src = tmp_path / "probe.cpp"
    src.write_text('#include "vkpbrt/vk_interop.hpp"\n#if VKPBRT_HAVE_VULKAN\n#error "expected no Vulkan"\n#endif\nint main() { return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", str(ROOT / "include"), str(src)], capture_output=True, text=True)
    if Path("/usr/include/vulkan/vulkan_core.h").exists():
        pytest.skip("this machine has system Vulkan headers")
    assert r.returncode == 0, r.stderr
