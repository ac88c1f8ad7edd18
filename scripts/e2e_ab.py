#!/usr/bin/env python
"""End-to-end (numpy in / numpy out, pinned host buffers) timing of pee_embed_batch + pee_extract_batch
for one workload; prints min / median ms over REPS repetitions.  Used to compare host-pipeline settings
(PEEB_PIPE_ROLES, PEEB_CHUNK_MB, PEEB_CHUNK_FIRST_MB), one process per setting.

    python scripts/e2e_ab.py [ct512|dx3000] [reps]
"""
import os
import statistics
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS  # noqa: E402
from codec_tcc_b200 import _cabi, pee  # noqa: E402
from codec_tcc_b200.synth import synth_batch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ct512"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n, h, w, maxval, bd, T = WORKLOADS[name]
imgs = _cabi.pinned_empty((n, h, w), np.uint16 if maxval > 255 else np.uint8)
imgs[...] = synth_batch(n, h, w, maxval, 2)
pay = _cabi.pinned_empty((n, (h * w + 7) // 8), np.uint8)
pay[...] = np.random.default_rng(7).integers(0, 256, pay.shape, dtype=np.uint8)
marked = _cabi.pinned_empty((n, h, w), imgs.dtype)
lm = _cabi.pinned_empty((n, h, (w + 7) // 8), np.uint8)
rec = _cabi.pinned_empty((n, h, w), imgs.dtype)
out = _cabi.pinned_empty((n, (h * w + 7) // 8), np.uint8)
_, _, info = pee.pee_embed_batch(imgs, pay, np.zeros(n, np.int64), T, bd, out_marked=marked, out_lm=lm)
cap = info[:, 2].astype(np.int64)


def step():
    t0 = time.perf_counter()
    pee.pee_embed_batch(imgs, pay, cap, T, bd, out_marked=marked, out_lm=lm)
    t1 = time.perf_counter()
    pee.pee_extract_batch(marked, lm, T, cap, bd, out_recovered=rec, out_payload=out)
    t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3


for _ in range(3):
    step()
ts = [step() for _ in range(reps)]
assert np.array_equal(rec, imgs)
tot = [a + b for a, b in ts]
npx = n * h * w
cfg = {k: v for k, v in os.environ.items() if k.startswith("PEEB_")}
print(f"{name} {cfg}: embed min {min(a for a, _ in ts):.2f} ms, extract min {min(b for _, b in ts):.2f} ms, "
      f"total min {min(tot):.2f} median {statistics.median(tot):.2f} ms -> {npx / min(tot) / 1e3:.0f} / "
      f"{npx / statistics.median(tot) / 1e3:.0f} Mpixel/s", flush=True)
