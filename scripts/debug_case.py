"""Development aid (GPU box): one embed/extract case against the C oracle with a report of where the arrays differ.
usage: python scripts/debug_case.py h w bit_depth itemsize T saturated seed [n_bits]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codec_tcc_b200 import pee  # noqa: E402
from codec_tcc_b200.synth import random_payload, synth_image, synth_saturated  # noqa: E402
from oracle import pee_c as PC  # noqa: E402

h, w, bd, itemsize, T, sat, seed = (int(x) for x in sys.argv[1:8])
maxval = (1 << bd) - 1
img = (synth_saturated if sat else synth_image)(h, w, maxval, seed)
if itemsize == 2 and img.dtype == np.uint8:
    img = img.astype(np.uint16)
_, _, i0 = PC.embed(img, np.zeros(img.size // 8 + 8, np.uint8), 0, T, bd)
n_bits = int(sys.argv[8]) if len(sys.argv) > 8 else int(i0["capacity"] * 0.6)
pay = random_payload(n_bits, int(sys.argv[9]) if len(sys.argv) > 9 else seed + 1)
m0, lm0, i0 = PC.embed(img, pay if pay.size else np.zeros(1, np.uint8), n_bits, T, bd)
i0.pop("status")
m1, lm1, i1 = pee.pee_embed(img, pay, T, bd, n_bits=n_bits)
print("info gpu", i1, "\ninfo ref", i0)
dm = np.argwhere(m1 != m0)
print("marked diffs", len(dm), "first", dm[:8].tolist(), "rows", sorted(set(dm[:, 0].tolist()))[:20] if len(dm) else [],
      "col range", (int(dm[:, 1].min()), int(dm[:, 1].max())) if len(dm) else None)
dl = np.argwhere(lm1 != lm0)
print("lm diffs", len(dl), dl[:8].tolist())
p1, r1 = pee.pee_extract(m0, lm0, T, n_bits, bd)
dr = np.argwhere(r1 != img)
print("extract (of the reference's marked image): recovered diffs", len(dr), dr[:8].tolist(), "payload equal", np.array_equal(p1, pay))
