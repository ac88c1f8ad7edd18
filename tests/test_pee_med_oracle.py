"""N1, CPU side: the causal (MED) PEE oracles -- the scalar C walk against the
vectorised numpy embed, and extract(embed(x)) == x with the payload read back, on random,
saturated, tiny and odd-shaped images.  *** parity unpinned *** (no PEE code in the reference)."""
import numpy as np
import pytest

from codec_tcc_b200.synth import random_payload, synth_image, synth_saturated
from oracle import pee_c, pee_numpy as PN

CASES = [
    (synth_image(64, 80, 4095, 1), 12, 4), (synth_image(33, 47, 65535, 2), 16, 96), (synth_image(50, 21, 255, 3), 8, 2),
    (synth_saturated(40, 56, 4095, 4), 12, 8), (synth_saturated(31, 33, 255, 5), 8, 1), (np.zeros((9, 12), np.uint16), 16, 3),
    (np.full((7, 7), 255, np.uint8), 8, 5), (synth_image(2, 9, 255, 6), 8, 2), (synth_image(9, 2, 4095, 7), 12, 2),
    (synth_image(1, 5, 255, 8), 8, 1), (synth_image(130, 257, 65535, 9), 16, 32768), (synth_image(20, 300, 1023, 10), 10, 512),
]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_med_c_vs_numpy_and_round_trip(idx):
    img, bd, T = CASES[idx]
    pay = random_payload(img.size, 100 + idx)
    m0, lm0, i0 = pee_c.embed(img, pay, img.size, T, bd, predictor="med")   # capacity probe
    cap = i0["capacity"]
    for n_bits in sorted({0, min(1, cap), cap // 2, cap}):
        m, lm, info = pee_c.embed(img, pay, n_bits, T, bd, predictor="med")
        m2, lm2, info2 = PN.med_embed(img, pay, T, bd, n_bits)
        assert np.array_equal(m, m2) and np.array_equal(lm, lm2) and info == info2
        assert info["status"] == 0 and info["capacity"] == cap and info["cap1"] == 0
        assert int(((m.astype(np.int64) - img.astype(np.int64)) ** 2).sum()) == info["sse"]
        assert np.array_equal(m[0], img[0]) and np.array_equal(m[:, 0], img[:, 0])        # border row / column untouched
        out, rec = pee_c.extract(m, lm, T, n_bits, predictor="med")
        assert np.array_equal(rec, img)
        assert np.array_equal(np.unpackbits(out)[:n_bits], np.unpackbits(pay)[:n_bits])
    if cap:
        with pytest.raises(ValueError):
            pee_c.extract(m0, lm0, T, cap + 1, predictor="med")
    assert pee_c.embed(img, pay, cap + 1, T, bd, predictor="med")[2]["status"] == -2
