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


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(INCLUDE, "peeb200.h"))
    return hs


def _compile_all(jobs, verbose: bool) -> None:
    """nvcc runs of independent translation units, side by side (peeb_pee2.cu alone takes minutes)."""
    from concurrent.futures import ThreadPoolExecutor

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1, 8)) as pool:
            list(pool.map(run, jobs))


BOUNDS_LIBPATH = os.path.join(LIBDIR, "bounds", "libpeeb200.so")
BOUNDS_SOURCE = "peeb_pee2.cu"


def build(force: bool = False, verbose: bool = False, with_bounds: bool = False) -> str:
    """-> path of libpeeb200.so.  with_bounds: also the bounds-checked build of the PEE band kernels
    (-DPEEB_DEBUG_BOUNDS, see peeb_pee2.cu) as lib/bounds/libpeeb200.so -- the same objects except peeb_pee2; it is
    loaded only by tests/test_gpu_bounds_build.py, through PEEB_LIBRARY."""
    os.makedirs(LIBDIR, exist_ok=True)
    headers = _headers()
    nvcc = _nvcc()
    extra = os.environ.get("PEEB_NVCC_EXTRA", "").split()
    objs, jobs = [], []
    for src in SOURCES:
        spath = os.path.join(CSRC, src)
        opath = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(opath)
        if force or _stale(opath, [spath] + headers):
            cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", INCLUDE, "-c", spath, "-o", opath]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)
    bobj = os.path.join(LIBDIR, "bounds", BOUNDS_SOURCE.replace(".cu", ".o"))
    if with_bounds:
        os.makedirs(os.path.dirname(BOUNDS_LIBPATH), exist_ok=True)
        spath = os.path.join(CSRC, BOUNDS_SOURCE)
        if force or _stale(bobj, [spath] + headers):
            jobs.append([nvcc, *NVCC_FLAGS, "-DPEEB_DEBUG_BOUNDS", "-I", INCLUDE, "-c", spath, "-o", bobj])
    _compile_all(jobs, verbose)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"]
    if force or _stale(LIBPATH, objs):
        subprocess.run([*link, "-o", LIBPATH, *objs], check=True)
    if with_bounds:
        bobjs = [bobj if o.endswith(BOUNDS_SOURCE.replace(".cu", ".o")) else o for o in objs]
        if force or _stale(BOUNDS_LIBPATH, bobjs):
            subprocess.run([*link, "-o", BOUNDS_LIBPATH, *bobjs], check=True)
    return LIBPATH


def build_bounds(force: bool = False) -> str:
    build(force=force, with_bounds=True)
    return BOUNDS_LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, with_bounds="--bounds" in sys.argv))
    if "--bounds" in sys.argv:
        print(BOUNDS_LIBPATH)
