"""N4, CPU side: the numpy restatement of the true inverse of the LSB path
(oracle/codec_numpy.py: recover_cover, extract_message_bits) against the embedders' own
outputs, including messages shorter than the segment plan (negative planned sizes,
src/codec.py:253-259) and segments longer than the image."""
import numpy as np
import pytest

from codec_tcc_b200.synth import synth_image, synth_saturated
from oracle import codec_numpy as OC


def _bits(n, seed):
    rng = np.random.default_rng(seed)
    return "".join("1" if b else "0" for b in rng.integers(0, 2, n).tolist())


@pytest.mark.parametrize("idx", range(3))
def test_inverse_of_both_embedders(idx):
    img, beta, sbs = [(synth_image(67, 45, 255, 11), 0.5, 8), (synth_image(40, 90, 4095, 12), 0.7, 16),
                      (synth_saturated(24, 56, 65535, 14), 0.9, 4)][idx]
    g, l = OC.adaptive_modalities_decomposition(img, beta=beta)
    s = len(l)
    for n in (0, 1, 2, 3, 7, 500, img.size, 3 * img.size):
        bits = _bits(n, n + idx)
        sizes, order = OC.segment_plan(s, len(bits))
        for align in (False, True):
            sp, bm, used, lens, order = OC.lsb_embed_block_then_multiplane(l, bits, search_block_size=sbs, align_across_planes=align)
            stego = OC.merge_modalities(g, sp).astype(img.dtype)
            meta = {"s": s, "segments_indices": order, "segments_lengths": lens, "hybrid": True, "align_across_planes": align,
                    "start_offset": OC.hybrid_start_offset(l[0], sbs), "message_bits": len(bits)}
            got = OC.extract_message_bits(stego, meta)
            assert len(got) == used
            if used == len(bits) and len(bits) >= 4 * s:
                assert got == bits
            assert np.array_equal(OC.recover_cover(stego, bm), img)
        sp, bm, used, lens, order = OC.lsb_embed_multi_plane(l, bits)
        stego = OC.merge_modalities(g, sp).astype(img.dtype)
        got = OC.extract_message_bits(stego, {"s": s, "segments_indices": order, "segments_lengths": lens})
        assert len(got) == used and np.array_equal(OC.recover_cover(stego, bm), img)
        if used == len(bits) and len(bits) >= 4 * s:
            assert got == bits


def test_text_round_trip():
    img = synth_image(64, 64, 4095, 3)
    g, l = OC.adaptive_modalities_decomposition(img, beta=0.8)
    msg = "Mensagem de teste para esteganografia!"  # src/codec.py:863
    sp, bm, used, lens, order = OC.lsb_embed_block_then_multiplane(l, OC.message_to_bits(msg), search_block_size=16)
    stego = OC.merge_modalities(g, sp).astype(img.dtype)
    meta = {"s": len(l), "segments_indices": order, "segments_lengths": lens, "hybrid": True,
            "start_offset": OC.hybrid_start_offset(l[0], 16), "message_bits": 8 * len(msg)}
    assert OC.extract_message(stego, meta) == msg          # the reference's decode_message returns garbage here
