"""The PEE parity suite once more, against the bounds-checked build of the band kernels.

compute-sanitizer cannot be used on the boxes this code is developed on, and the sweep reads a word before and
after each cell on purpose (into slack the shared layout leaves for it).  `python -m codec_tcc_b200.build --bounds`
compiles peeb_pee2.cu with -DPEEB_DEBUG_BOUNDS: every index the kernels derive from the geometry is compared with
its region's size on the device (see the BOUNDS sites in the source), violations are counted, not trapped.  The
suite runs in a child process with PEEB_LIBRARY pointing at that build; afterwards the counters must show that
checks ran and none failed."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = ["tests/test_gpu_pee.py", "tests/test_gpu_pee_random.py"]


def _bounds_library():
    from codec_tcc_b200 import build as B

    if not os.path.exists(B.BOUNDS_LIBPATH):
        B.build_bounds()  # nvcc is part of the image; normally __graft_entry__.build() has done this already
    return B.BOUNDS_LIBPATH


@pytest.mark.gpu
def test_checker_reports_a_violation_when_there_is_one(tmp_path):
    code = (
        "import json, sys\n"
        "from codec_tcc_b200 import _cabi\n"
        "import ctypes as C\n"
        "out = (C.c_ulonglong * 6)()\n"
        "_cabi.check(_cabi.lib().peeb_debug_bounds(out, 2), 'selftest')\n"
        "print(json.dumps([int(v) for v in out]))\n"
    )
    env = dict(os.environ, PEEB_LIBRARY=_bounds_library())
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    violations, site, off, lim, _items, checking = json.loads(r.stdout.strip().splitlines()[-1])
    assert checking == 1
    assert (violations, site, off, lim) == (1, 99, 16, 16)


@pytest.mark.gpu
def test_parity_suite_on_the_bounds_checked_build(tmp_path):
    report = tmp_path / "bounds.json"
    env = dict(os.environ, PEEB_LIBRARY=_bounds_library(), PEEB_BOUNDS_REPORT=str(report))
    r = subprocess.run([sys.executable, "-m", "pytest", *SUITE, "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-1000:])
    rep = json.loads(report.read_text())
    assert rep["checking"], "the child did not load the bounds-checked build"
    assert rep["items_checked"] > 100000, rep
    assert rep["violations"] == 0, (rep, r.stderr[-1500:])


def test_product_build_does_not_check():
    """The normal library exports the call and says that it checks nothing (no GPU work: all zeros)."""
    import ctypes as C

    from codec_tcc_b200 import _cabi

    if os.environ.get("PEEB_LIBRARY"):
        pytest.skip("running against an alternative build")
    out = (C.c_ulonglong * 6)(*([7] * 6))
    assert _cabi.lib().peeb_debug_bounds(out, 0) == 0
    assert [int(v) for v in out] == [0] * 6
