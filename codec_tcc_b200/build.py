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


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
