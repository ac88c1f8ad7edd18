import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "reference_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_images():
    """The six images the golden file was generated from (two reference
    fixtures from fixtures.npz + four seeded synthetic ones)."""
    from codec_tcc_b200.synth import synth_image, synth_saturated

    z = np.load(os.path.join(GOLDEN_DIR, "fixtures.npz"))
    return {
        "pe": z["pe"], "torax": z["torax"],
        "synth16_257x301": synth_image(257, 301, 65535, 11),
        "synth12_300x200": synth_image(300, 200, 4095, 12),
        "synth8_129x70": synth_image(129, 70, 255, 13),
        "sat12_96x160": synth_saturated(96, 160, 4095, 14),
    }


def pytest_sessionfinish(session, exitstatus):
    """A run against the bounds-checked build (tests/test_gpu_bounds_build.py starts one with
    PEEB_LIBRARY / PEEB_BOUNDS_REPORT set) leaves the library's counters behind for its parent."""
    path = os.environ.get("PEEB_BOUNDS_REPORT")
    if not path:
        return
    from codec_tcc_b200 import _cabi

    rep = _cabi.debug_bounds()
    rep["exitstatus"] = int(exitstatus)
    with open(path, "w") as f:
        json.dump(rep, f)
