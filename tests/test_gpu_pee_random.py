"""GPU: randomised PEE parity against the C oracle -- shapes from 3x3 to wide rows that need the
512- and 1024-thread band kernels, every bit depth, random thresholds and payload lengths."""
import numpy as np
import pytest

from codec_tcc_b200 import pee
from codec_tcc_b200.synth import random_payload, synth_image, synth_saturated

from oracle import pee_c as PC

pytestmark = pytest.mark.gpu


def _one(rng, h, w, bd, itemsize):
    maxval = (1 << bd) - 1
    gen = synth_saturated if rng.integers(0, 2) else synth_image
    img = gen(h, w, maxval, int(rng.integers(0, 1 << 30)))
    if itemsize == 2 and img.dtype == np.uint8:
        img = img.astype(np.uint16)
    T = int(rng.integers(1, min(1 << (bd - 1), 64) + 1))
    _, _, i0 = PC.embed(img, np.zeros(img.size // 8 + 8, np.uint8), 0, T, bd)
    cap = i0["capacity"]
    n_bits = int(rng.integers(0, max(1, int(cap * 0.95)) + 1)) if cap else 0
    pay = random_payload(n_bits, int(rng.integers(0, 1 << 30)))
    m0, lm0, i0 = PC.embed(img, pay if pay.size else np.zeros(1, np.uint8), n_bits, T, bd)
    if i0.pop("status") != 0:
        return
    m1, lm1, i1 = pee.pee_embed(img, pay, T, bd, n_bits=n_bits)
    assert i1 == i0, (h, w, bd, T, n_bits, i1, i0)
    assert np.array_equal(m1, m0) and np.array_equal(lm1, lm0), (h, w, bd, T, n_bits)
    p1, r1 = pee.pee_extract(m1, lm1, T, n_bits, bd)
    assert np.array_equal(r1, img) and np.array_equal(p1, pay), (h, w, bd, T, n_bits)


@pytest.fixture(params=["bands", "cluster"])
def kernel_path(request, monkeypatch):
    monkeypatch.setenv("PEEB_CLUSTER", "1" if request.param == "cluster" else "0")
    return request.param


@pytest.mark.parametrize("seed", range(6))
def test_random_small_shapes(seed, kernel_path):
    rng = np.random.default_rng(seed)
    for _ in range(25):
        h, w = int(rng.integers(3, 90)), int(rng.integers(3, 400))
        bd, itemsize = [(8, 1), (8, 2), (10, 2), (12, 2), (16, 2), (5, 1)][int(rng.integers(0, 6))]
        _one(rng, h, w, bd, itemsize)


@pytest.mark.parametrize("w", [1024, 2048, 3000, 4096, 5000, 8192, 1000, 6001])
def test_wide_rows(w, kernel_path):
    rng = np.random.default_rng(w)
    for bd, itemsize in ((12, 2), (8, 1)):
        _one(rng, int(rng.integers(20, 70)), w, bd, itemsize)
