"""Builds libpeeb200.so (hand-written sm_100a CUDA behind the C ABI of
include/peeb200.h) with nvcc, in tree.  nvcc cross-compiles without a GPU."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")
LIBDIR = os.path.join(PKG, "lib")
LIBPATH = os.path.join(LIBDIR, "libpeeb200.so")
SOURCES = ["peeb_api.cu", "peeb_moments.cu", "peeb_lsb.cu", "peeb_pee.cu", "peeb_pee2.cu", "peeb_pee_med.cu", "peeb_bitcode.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libpeeb200 cannot be built")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "peeb200.h"))
    nvcc = _nvcc()
    extra = os.environ.get("PEEB_NVCC_EXTRA", "").split()
    objs = []
    for src in SOURCES:
        spath = os.path.join(CSRC, src)
        opath = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(opath)
        if force or _stale(opath, [spath] + headers):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", INCLUDE, "-c", spath, "-o", opath]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _stale(LIBPATH, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIBPATH, *objs]
        subprocess.run(cmd, check=True)
    return LIBPATH


BOUNDS_LIBPATH = os.path.join(LIBDIR, "bounds", "libpeeb200.so")


def build_bounds(force: bool = False) -> str:
    """The bounds-checked build of the PEE band kernels (-DPEEB_DEBUG_BOUNDS, see peeb_pee2.cu): same objects as
    the product library except peeb_pee2.  Loaded only by tests/test_gpu_bounds_build.py through PEEB_LIBRARY."""
    build(force=False)
    os.makedirs(os.path.dirname(BOUNDS_LIBPATH), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "peeb200.h"))
    nvcc = _nvcc()
    spath = os.path.join(CSRC, "peeb_pee2.cu")
    opath = os.path.join(LIBDIR, "bounds", "peeb_pee2.o")
    if force or _stale(opath, [spath] + headers):
        subprocess.run([nvcc, *NVCC_FLAGS, "-DPEEB_DEBUG_BOUNDS", "-I", INCLUDE, "-c", spath, "-o", opath], check=True)
    objs = [opath if src == "peeb_pee2.cu" else os.path.join(LIBDIR, src.replace(".cu", ".o")) for src in SOURCES]
    if force or _stale(BOUNDS_LIBPATH, objs):
        subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", BOUNDS_LIBPATH, *objs], check=True)
    return BOUNDS_LIBPATH


if __name__ == "__main__":
    if "--bounds" in sys.argv:
        print(build_bounds(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
