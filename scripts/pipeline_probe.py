"""Development aid: wall time of codec.embed_pipeline (the reference's encode flow, device resident) per call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from codec_tcc_b200 import codec
from codec_tcc_b200.synth import synth_image

img = synth_image(3000, 3000, 4095, 50)
rng = np.random.default_rng(1)
bits = "".join("1" if b else "0" for b in rng.integers(0, 2, 18_000_000).tolist())
for k in range(6):
    t0 = time.perf_counter()
    st, bm, meta = codec.embed_pipeline(img, bits, beta=0.8, search_block_size=16)
    print(f"call {k}: {(time.perf_counter() - t0) * 1e3:.1f} ms  s={meta['s']}", flush=True)
