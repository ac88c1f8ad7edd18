"""CPU, build container only: the numpy restatements against the UNMODIFIED
reference executed live (skipped where /root/reference is absent, e.g. on the
GPU box -- the committed golden vectors cover that case)."""
import numpy as np
import pytest

from oracle import codec_numpy as OC
from oracle import mse_numpy as OM
from oracle import ref_import

from codec_tcc_b200.synth import synth_image, synth_saturated

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")


def _bits(n, seed):
    rng = np.random.default_rng(seed)
    return "".join("1" if b else "0" for b in rng.integers(0, 2, n).tolist())


CASES = [
    (synth_image(67, 45, 255, 1), 0.5, 8),
    (synth_image(50, 130, 4095, 2), 0.7, 16),
    (synth_image(33, 33, 65535, 3), 0.9, 7),
    (synth_saturated(40, 56, 255, 4), 0.3, 4),
    (np.zeros((20, 24), np.uint16), 0.8, 8),          # constant image
    (np.full((9, 9), 255, np.uint8), 0.8, 16),        # constant, tile larger than image
]


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_lsb_path_matches_reference(idx):
    img, beta, sbs = CASES[idx]
    R = ref_import.codec()
    with ref_import.quiet():
        g0, l0 = R.adaptive_modalities_decomposition(img, beta=beta)
    g1, l1 = OC.adaptive_modalities_decomposition(img, beta=beta)
    assert len(l0) == len(l1)
    assert all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip(l0 + g0, l1 + g1))
    for nbits_payload in (0, 1, 5, 1000, img.size * 3):
        bits = _bits(nbits_payload, idx)
        for align in (False, True):
            r = R.lsb_embed_block_then_multiplane(l0, bits, search_block_size=sbs, align_across_planes=align)
            o = OC.lsb_embed_block_then_multiplane(l1, bits, search_block_size=sbs, align_across_planes=align)
            _same_embed(r, o)
        r = R.lsb_embed_multi_plane(l0, bits)
        o = OC.lsb_embed_multi_plane(l1, bits)
        _same_embed(r, o)
        meta = {"s": len(l0), "segments_indices": r[4], "segments_lengths": r[3]}
        assert R.decode_message(r[0], [b.ravel() for b in r[1]], meta) == \
            OC.decode_message(o[0], [b.ravel() for b in o[1]], meta)
        assert np.array_equal(R.merge_modalities(g0, r[0]), OC.merge_modalities(g1, o[0]))


def _same_embed(r, o):
    assert r[2] == o[2] and list(r[3]) == list(o[3]) and list(r[4]) == list(o[4])
    for a, b in zip(r[0], o[0]):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    for a, b in zip(r[1], o[1]):
        assert a.dtype == b.dtype == np.uint8 and np.array_equal(a, b)


def test_entropy_mi_random():
    R = ref_import.codec()
    for seed in range(4):
        img = synth_image(40 + seed, 50, [255, 4095, 65535, 1023][seed], seed)
        assert float(R.calculate_entropy(img)) == float(OC.calculate_entropy(img))
        for i in range(8 * img.dtype.itemsize):
            pl = (img >> i) & 1
            assert float(R.calculate_mutual_information(pl, img)) == float(OC.calculate_mutual_information(pl, img))


def test_general_mutual_information_random():
    """planes that are not bit planes of the image: the restatement's joint-bincount branch, live"""
    R = ref_import.codec()
    rng = np.random.default_rng(3)
    for seed in range(4):
        img = synth_image(40 + seed, 50, [255, 4095, 65535, 1023][seed], seed)
        other = synth_image(40 + seed, 50, [255, 4095, 65535, 1023][seed], seed + 100)
        for plane in ((other >> 2) & 1, rng.integers(0, 2, img.shape).astype(img.dtype), (other % 5).astype(np.uint8)):
            assert float(R.calculate_mutual_information(plane, img)) == float(OC.calculate_mutual_information(plane, img))


def test_metrics_match_reference():
    an = ref_import.mse().AnalisadorMSE()
    rng = np.random.default_rng(0)
    for maxval in (255, 4095, 65535):
        a = synth_image(64, 48, maxval, 5)
        b = a.copy()
        b[rng.integers(0, 64, 50), rng.integers(0, 48, 50)] ^= 1
        c = a.copy(); c[3, 3] = min(maxval, int(a.max()) + 7) if a.max() < maxval else a.max() - 1
        for x, y in ((a, b), (a, a), (a, c), (c, b)):
            with ref_import.quiet():
                m0, r0 = an.calcular_mse(x, y)
                s0 = an.calcular_ssim_simples(x, y)
            m1, r1 = OM.calcular_mse(x, y)
            assert float(m0) == float(m1) and float(r0) == float(r1)
            assert float(s0) == float(OM.calcular_ssim_simples(x, y))
            assert float(an.calcular_psnr(m0, r0)) == float(OM.calcular_psnr(m1, r1))
