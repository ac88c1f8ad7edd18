"""Small PEE / LSB / metrics run for compute-sanitizer (memcheck): odd sizes, partial strips,
both staging paths."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
from codec_tcc_b200 import _cabi, codec, mse, pee
from codec_tcc_b200.synth import random_payload, synth_image, synth_saturated

for bulk in (True, False):
    _cabi.workspace().set_option("bulk", bulk)
    for (h, w, mv, bd) in ((70, 131, 4095, 12), (64, 512, 65535, 16), (33, 200, 255, 8), (5, 7, 255, 8)):
        img = synth_saturated(h, w, mv, 3)
        pay = random_payload(200, 1)
        m, lm, info = pee.pee_embed(img, pay, 3, bd, n_bits=0)
        out, rec = pee.pee_extract(m, lm, 3, 0, bd)
        assert np.array_equal(rec, img)
        rows = pee.pee_sweep(img, random_payload(img.size, 2), [1, 4], bd, n_bits=img.size)
        cap = rows[1]["capacity"]
        if cap:
            m, lm, info = pee.pee_embed(img, random_payload(cap // 2, 5), 4, bd, n_bits=cap // 2)
            out, rec = pee.pee_extract(m, lm, 4, cap // 2, bd)
            assert np.array_equal(rec, img)
        pee.pee_histogram(img, bd)
img = synth_image(129, 70, 4095, 1)
g, l = codec.adaptive_modalities_decomposition(img, beta=0.5)
st, bm, used, lens, idx = codec.lsb_embed_block_then_multiplane(l, codec.message_to_bits("sanitize me"), 16)
stego = codec.merge_modalities(g, st)
codec.decode_message(codec.extract_local_planes(stego, len(l)), bm, {"s": len(l), "segments_indices": idx, "segments_lengths": lens})
print(mse.AnalisadorMSE().calcular_mse(img, stego))
print("sanitize smoke ok")
