"""Out-of-bounds accesses of the kernels on the emulator: HOSTSIM_GUARD=end|start (tests/hostsim/hostsim.cpp) places every
device allocation against an inaccessible page, so a kernel that reads or writes past a plane -- a vector load across
the end of a ragged row, a gather with a bad coordinate, an apron row that does not exist -- faults instead of
touching a neighbouring allocation.  The emulator's stand-in for compute-sanitizer memcheck (which covers the
GPU-confirmed sizes: profiles/r02_sanitizer.txt), applied to the seeded size fuzz and the ragged parity cases.  The guard
mode is fixed when the emulator makes its first allocation, hence the subprocesses."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]

PROBE = """
import ctypes, sys
sys.path.insert(0, {root!r})
from vulkanpbrt_b200 import _capi
_capi._lib = _capi.configure(ctypes.CDLL({lib!r}))
from vulkanpbrt_b200 import Context, DescriptorImage
img = DescriptorImage.create(Context(0), _capi.FORMAT_R8_UNORM, 3, 1)
img.compile()
p = img.info().data
assert ctypes.c_uint8.from_address(p + 2).value == 0
print("inside ok", flush=True)
ctypes.c_uint8.from_address(p + int(sys.argv[1])).value
print("outside readable", flush=True)
"""


@pytest.mark.parametrize("mode,offset", [("end", 16), ("start", -1)])
def test_guard_mode_faults_on_an_out_of_bounds_access(mode, offset):
    """the detector detects: one byte past the (16-byte granular) end / before the start of an allocation is a SIGSEGV"""
    subprocess.run(["make", "-C", str(ROOT / "tests" / "hostsim")], check=True, capture_output=True)
    code = PROBE.format(root=str(ROOT), lib=str(ROOT / "tests" / "hostsim" / "libvkpbrt_hostsim.so"))
    r = subprocess.run([sys.executable, "-c", code, str(offset)], capture_output=True, text=True, env=dict(os.environ, HOSTSIM_GUARD=mode))
    assert "inside ok" in r.stdout and "outside readable" not in r.stdout and r.returncode < 0, (r.returncode, r.stdout, r.stderr[-300:])
    r = subprocess.run([sys.executable, "-c", code, str(offset)], capture_output=True, text=True, env={k: v for k, v in os.environ.items() if k != "HOSTSIM_GUARD"})
    assert r.returncode == 0 and "outside readable" in r.stdout          # without the guard the same access goes unnoticed


@pytest.mark.timeout(900)
@pytest.mark.parametrize("mode,selection", [("end", ["tests/test_fuzz_sizes.py", "tests/test_parity.py", "-k", "random_sizes or negative_jitter or narrower or other_block or bfr_block or x8x16x32"]),
                                            ("start", ["tests/test_fuzz_sizes.py", "-k", "random_sizes"])])
def test_no_kernel_touches_memory_outside_its_planes(mode, selection):
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider"] + selection, cwd=str(ROOT), capture_output=True, text=True,
                       env=dict(os.environ, HOSTSIM_GUARD=mode))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout
