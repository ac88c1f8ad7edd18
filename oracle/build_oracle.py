"""Compile oracle/pee_ref.c -> oracle/libpee_oracle.so (gcc, -O2, OpenMP when
available).  Test infrastructure only."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pee_ref.c")
OUT = os.path.join(HERE, "libpee_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    base = ["gcc", "-O2", "-fPIC", "-shared", "-std=c11", "-Wall", "-o", OUT, SRC]
    try:
        subprocess.run(base[:1] + ["-fopenmp"] + base[1:], check=True, capture_output=True)
    except (subprocess.CalledProcessError, FileNotFoundError):
        subprocess.run(base, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
