"""N1, GPU parity: PEE with the causal MED predictor (parallel embed, anti-diagonal
wavefront extract) through the C ABI, bit for bit against the scalar C oracle, plus the round trip.
*** parity unpinned *** (the reference has no PEE code)."""
import numpy as np
import pytest

from codec_tcc_b200 import pee
from codec_tcc_b200.synth import random_payload, synth_batch, synth_image, synth_saturated
from oracle import pee_c

pytestmark = pytest.mark.gpu

CASES = [
    (synth_image(64, 80, 4095, 1), 12, 4), (synth_image(33, 47, 65535, 2), 16, 96), (synth_image(50, 21, 255, 3), 8, 2),
    (synth_saturated(40, 56, 4095, 4), 12, 8), (synth_saturated(31, 33, 255, 5), 8, 1), (np.zeros((9, 12), np.uint16), 16, 3),
    (np.full((7, 7), 255, np.uint8), 8, 5), (synth_image(2, 9, 255, 6), 8, 2), (synth_image(9, 2, 4095, 7), 12, 2),
    (synth_image(1, 5, 255, 8), 8, 1), (synth_image(130, 257, 65535, 9), 16, 32768), (synth_image(20, 300, 1023, 10), 10, 512),
    (synth_image(512, 512, 65535, 11), 16, 96), (synth_image(700, 1030, 4095, 12), 12, 12), (synth_image(67, 1000, 255, 13), 8, 3),
]


@pytest.fixture(params=["one_cta", "cluster_of_two"])
def extract_form(request, monkeypatch):
    """The wavefront extract as one CTA per image and as a cluster of two CTAs per image (line buffers handed over
    through distributed shared memory); PEEB_MED_CLUSTER forces the form."""
    monkeypatch.setenv("PEEB_MED_CLUSTER", "2" if request.param == "cluster_of_two" else "1")
    return request.param


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_med_gpu_vs_oracle(idx, extract_form):
    img, bd, T = CASES[idx]
    pay = random_payload(img.size, 200 + idx)
    cap = pee_c.embed(img, pay, img.size, T, bd, predictor="med")[2]["capacity"]
    for n_bits in sorted({0, min(1, cap), cap // 3, cap}):
        m0, lm0, i0 = pee_c.embed(img, pay, n_bits, T, bd, predictor="med")
        m, lm, info = pee.pee_embed(img, pay, T, bd, n_bits=n_bits, predictor="med")
        assert np.array_equal(m, m0) and np.array_equal(lm, lm0)
        assert {k: info[k] for k in ("capacity", "cap0", "cap1", "n_flagged", "sse")} == {k: i0[k] for k in ("capacity", "cap0", "cap1", "n_flagged", "sse")}
        out, rec = pee.pee_extract(m, lm, T, n_bits, bd, predictor="med")
        assert np.array_equal(rec, img)
        assert np.array_equal(np.unpackbits(out)[:n_bits], np.unpackbits(pay)[:n_bits])
    with pytest.raises(ValueError):
        pee.pee_embed(img, pay, T, bd, n_bits=cap + 1, predictor="med")


@pytest.mark.parametrize("rows", [3, 16])
def test_med_embed_multi_row_items(rows, monkeypatch):
    """The embed kernels walk `rows` consecutive rows per warp item (the library picks 1 for small batches,
    up to 16 for large ones): same bits for every choice, also when h is not a multiple of it."""
    monkeypatch.setenv("PEEB_MED_ROWS", str(rows))
    for idx in (0, 1, 3, 4, 10, 11, 13, 14):
        img, bd, T = CASES[idx]
        pay = random_payload(img.size, 300 + idx)
        cap = pee_c.embed(img, pay, img.size, T, bd, predictor="med")[2]["capacity"]
        m0, lm0, i0 = pee_c.embed(img, pay, cap, T, bd, predictor="med")
        m, lm, info = pee.pee_embed(img, pay, T, bd, n_bits=cap, predictor="med")
        assert np.array_equal(m, m0) and np.array_equal(lm, lm0), (rows, idx)
        assert {k: info[k] for k in ("capacity", "n_flagged", "sse")} == {k: i0[k] for k in ("capacity", "n_flagged", "sse")}
        out, rec = pee.pee_extract(m, lm, T, cap, bd, predictor="med")
        assert np.array_equal(rec, img) and np.array_equal(np.unpackbits(out)[:cap], np.unpackbits(pay)[:cap])


def test_med_batch_and_auto_threshold(extract_form):
    imgs = synth_batch(5, 96, 160, 4095, 31)
    pays = np.stack([random_payload(imgs[0].size, 70 + k) for k in range(5)])
    nb = np.array([0, 100, 2000, 2900, 50], np.int64)
    marked, lm, info = pee.pee_embed_batch(imgs, pays, nb, 6, 12, predictor="med")
    for u in range(5):
        m0, lm0, i0 = pee_c.embed(imgs[u], pays[u], int(nb[u]), 6, 12, predictor="med")
        assert np.array_equal(marked[u], m0) and np.array_equal(lm[u], lm0) and int(info[u, 6]) == i0["sse"] and int(info[u, 7]) == i0["status"]
    out, rec, xinfo = pee.pee_extract_batch(marked, lm, 6, nb, 12, predictor="med")
    assert np.array_equal(rec, imgs)
    for u in range(5):
        assert np.array_equal(np.unpackbits(out[u])[:nb[u]], np.unpackbits(pays[u])[:nb[u]])
    img = synth_image(128, 128, 4095, 77)
    bits = random_payload(4000, 5)
    m, lmp, info = pee.pee_embed(img, bits, None, 12, n_bits=4000, predictor="med")     # smallest T that fits
    assert info["capacity"] >= 4000 and (info["T"] == 1 or pee_c.embed(img, bits, 4000, info["T"] - 1, 12, predictor="med")[2]["status"] == -2)
    out, rec = pee.pee_extract(m, lmp, info["T"], 4000, 12, predictor="med")
    assert np.array_equal(rec, img) and np.array_equal(np.unpackbits(out)[:4000], np.unpackbits(bits)[:4000])


def test_med_batch_threshold_search_on_the_device():
    """T=None for a batch with the causal predictor: every unit gets the smallest T (from 1 upwards) whose capacity
    holds its payload; same T, marked image and location map as the oracle run at that T, smaller T do not fit."""
    imgs = synth_batch(6, 80, 120, 4095, 77)
    pays = np.stack([random_payload(imgs[0].size, 500 + k) for k in range(6)])
    caps = [pee_c.embed(imgs[u], pays[u], imgs[0].size, 2048, 12, predictor="med")[2]["capacity"] for u in range(6)]
    nb = np.array([0, 30, int(caps[2] * 0.2), int(caps[3] * 0.5), int(caps[4] * 0.8), imgs[0].size], np.int64)
    marked, lm, info = pee.pee_embed_batch(imgs, pays, nb, None, 12, predictor="med")
    for u in range(5):
        T = int(info[u, 0])
        assert int(info[u, 7]) == 0 and int(info[u, 2]) >= nb[u], u
        m0, lm0, i0 = pee_c.embed(imgs[u], pays[u], int(nb[u]), T, 12, predictor="med")
        assert i0["status"] == 0 and np.array_equal(marked[u], m0) and np.array_equal(lm[u], lm0), u
        if T > 1:
            assert pee_c.embed(imgs[u], pays[u], int(nb[u]), T - 1, 12, predictor="med")[2]["status"] != 0, u
    assert int(info[5, 7]) == pee.PEEB_E_CAPACITY and int(info[5, 0]) == 2048
    ok = info[:, 7] == 0
    out, rec, _ = pee.pee_extract_batch(marked[ok], lm[ok], info[ok, 0].astype(np.int32), nb[ok], 12, predictor="med")
    assert np.array_equal(rec, imgs[ok])
