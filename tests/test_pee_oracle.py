"""CPU: the two independent PEE oracles (numpy masks+cumsum vs scalar C raster
walk) against each other, and the properties the specification promises
(SURVEY.md Appendix A).  PEE parity is UNPINNED -- the reference has no PEE."""
import numpy as np
import pytest

from oracle import pee_c as PC
from oracle import pee_numpy as PN

from codec_tcc_b200.synth import random_payload, synth_image, synth_saturated

SHAPES = [(3, 3), (3, 40), (40, 3), (5, 7), (17, 129), (64, 64), (70, 131), (257, 301)]


def _capacity(img, T, bd):
    _, _, info = PC.embed(img, np.zeros(img.size // 8 + 8, np.uint8), 0, T, bd)
    return info["capacity"]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("maxval,bd", [(255, 8), (4095, 12), (65535, 16)])
def test_numpy_vs_c_and_roundtrip(shape, maxval, bd):
    h, w = shape
    for gen, seed in ((synth_image, 3), (synth_saturated, 4)):
        img = gen(h, w, maxval, seed)
        for T in (1, 3, 20):
            n_bits = int(_capacity(img, T, bd) * 0.9)
            pay = random_payload(n_bits, seed + T)
            m1, lm1, i1 = PN.pee_embed(img, pay, T, bd, n_bits=n_bits)
            m2, lm2, i2 = PC.embed(img, pay, n_bits, T, bd)
            assert i2.pop("status") == 0 and i1 == i2
            assert np.array_equal(m1, m2) and np.array_equal(lm1, lm2)
            assert m1.dtype == img.dtype and lm1.shape == (h, (w + 7) // 8)
            assert int(m1.max()) <= maxval
            # borders never change
            assert np.array_equal(m1[0], img[0]) and np.array_equal(m1[-1], img[-1])
            assert np.array_equal(m1[:, 0], img[:, 0]) and np.array_equal(m1[:, -1], img[:, -1])
            p1, r1 = PN.pee_extract(m1, lm1, T, n_bits, bd)
            p2, r2 = PC.extract(m2, lm2, T, n_bits)
            assert np.array_equal(r1, img) and np.array_equal(r2, img)
            assert np.array_equal(p1, pay) and np.array_equal(p2, pay)


def test_too_small_images_carry_nothing():
    for shape in ((1, 1), (2, 9), (9, 2), (1, 50)):
        img = synth_image(shape[0], shape[1], 255, 1)
        m, lm, info = PN.pee_embed(img, b"", 2, 8, n_bits=0)
        assert info["capacity"] == 0 and np.array_equal(m, img) and not lm.any()
        with pytest.raises(ValueError):
            PN.pee_embed(img, b"\xff", 2, 8, n_bits=1)


def test_overflow_rejected_and_reported():
    img = synth_image(32, 32, 255, 2)
    cap = _capacity(img, 2, 8)
    with pytest.raises(ValueError):
        PN.pee_embed(img, random_payload(cap + 500, 1), 2, 8, n_bits=cap + 500)
    _, _, info = PC.embed(img, random_payload(cap + 500, 1), cap + 500, 2, 8)
    assert info["status"] == -2


def test_histogram_and_auto_threshold():
    for maxval, bd in ((255, 8), (4095, 12)):
        img = synth_image(90, 70, maxval, 7)
        h_np = PN.error_histogram(img, bd)
        h_c = PC.hist(img, bd)
        assert np.array_equal(h_np, h_c)
        # pass-0 capacity from the histogram is exact (Appendix A)
        for T in (1, 2, 5):
            tmax = h_np.shape[1] // 2
            _, _, info = PC.embed(img, np.zeros(1, np.uint8), 0, T, bd)
            assert info["cap0"] == int(h_np[0, tmax - T:tmax + T].sum())
        n_bits = 2500
        pay = random_payload(n_bits, 9)
        m, lm, info = PN.pee_embed(img, pay, None, bd, n_bits=n_bits)
        assert info["capacity"] >= n_bits
        if info["T"] > 1:  # minimality of the estimate-then-verify rule
            assert PN.estimate_T(h_np, n_bits) <= info["T"]
        p, r = PN.pee_extract(m, lm, info["T"], n_bits, bd)
        assert np.array_equal(r, img) and np.array_equal(p, pay)


def test_survey_reference_point(golden_images):
    """SURVEY.md Appendix A quotes pe.dcm, B=12, T=1: 1 127 flagged pixels."""
    pe = golden_images["pe"]
    cap = _capacity(pe, 1, 12)
    bits = np.random.default_rng(0).integers(0, 2, cap).astype(np.uint8)
    _, _, info = PC.embed(pe, np.packbits(bits), cap, 1, 12)
    assert info["n_flagged"] == 1127 and info["status"] == 0


def test_sweep_monotone():
    img = synth_image(64, 64, 4095, 5)
    rows = PN.pee_sweep(img, random_payload(img.size, 1), [1, 2, 4, 8, 16], 12)
    caps = [r["capacity"] for r in rows]
    assert caps == sorted(caps)
    assert all(r["psnr"] > 30 for r in rows)


def test_batch_c():
    imgs = np.stack([synth_image(48, 40, 4095, s) for s in range(5)])
    nb = np.array([100, 0, 17, 150, 50], np.int64)
    pays = np.zeros((5, 80), np.uint8)
    for u in range(5):
        p = random_payload(int(nb[u]), u)
        pays[u, :p.size] = p
    marked, lm, info = PC.embed_batch(imgs, pays, nb, 3, 12, threads=2)
    out, rec, rc = PC.extract_batch(marked, lm, 3, nb, 80, threads=2)
    assert rc == 0 and np.array_equal(rec, imgs) and np.array_equal(out, pays)
    for u in range(5):
        m1, l1, i1 = PN.pee_embed(imgs[u], pays[u], 3, 12, n_bits=int(nb[u]))
        assert np.array_equal(m1, marked[u]) and np.array_equal(l1, lm[u]) and i1["sse"] == info[u, 6]
