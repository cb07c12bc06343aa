"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

    libvkpbrt_b200.so   CUDA kernels (sm_100a) + the C ABI of include/vkpbrt_b200.h
    libvkpbrt_synth.so  synthetic G-buffer sequence generator (host C, input tooling)

Usage: python -m vulkanpbrt_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
LIB = PKG / "lib"
OBJ = ROOT / "build" / "obj"

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOST_CC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-ccbin", HOST_CXX,
          "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-Wall"]

# per-file extra flags: the streaming kernels whose outputs must be bit-exact never contract FMAs
UNITS = {
    "accumulate.cu": ["-fmad=false"],
    "taa.cu": ["-fmad=false"],
    "bmfr.cu": [],
    "bfr.cu": [],
    "halo.cu": [],
    "debug.cu": [],
    "convert.cu": ["-fmad=false"],
    "api.cpp": [],
}


def _run(cmd, verbose):
    if verbose:
        print(" ".join(str(c) for c in cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {' '.join(str(c) for c in cmd)}")
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build_cuda(force=False, verbose=False) -> Path:
    LIB.mkdir(parents=True, exist_ok=True)
    OBJ.mkdir(parents=True, exist_ok=True)
    headers = (list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + list(CSRC.glob("*.inc")) + [ROOT / "include" / "vkpbrt_b200.h"]
               + list((ROOT / "include" / "vkpbrt").glob("*.hpp")))       # api.cpp includes the C++ layer (vkpbrt::BandedRank)
    objs = []
    for name, extra in UNITS.items():
        src = CSRC / name
        obj = OBJ / (name + ".o")
        if force or _stale(obj, [src] + headers):
            lang = ["-x", "cu"] if name.endswith(".cu") else []
            _run([NVCC, *COMMON, *ARCH, *extra, *lang, "-c", str(src), "-o", str(obj)], verbose)
        objs.append(obj)
    so = LIB / "libvkpbrt_b200.so"
    if force or _stale(so, objs):
        _run([NVCC, "-shared", "-ccbin", HOST_CXX, *ARCH, "-cudart", "static", "-o", str(so), *map(str, objs)], verbose)
    return so


def build_synth(force=False, verbose=False) -> Path:
    LIB.mkdir(parents=True, exist_ok=True)
    src = PKG / "synth" / "synth.c"
    so = LIB / "libvkpbrt_synth.so"
    if force or _stale(so, [src]):
        _run([HOST_CC, "-O2", "-std=c11", "-fPIC", "-fopenmp", "-fvisibility=hidden", "-shared", "-o", str(so), str(src), "-lm"],
             verbose)
    return so


def build_all(force=False, verbose=False):
    return build_cuda(force, verbose), build_synth(force, verbose)


if __name__ == "__main__":
    force = "--force" in sys.argv
    verbose = "--verbose" in sys.argv or "-v" in sys.argv
    for p in build_all(force, verbose):
        print(p)
