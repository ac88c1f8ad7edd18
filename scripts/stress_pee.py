"""Development aid (GPU box): many random shapes / depths / thresholds through both rhombus kernel families (band
kernels, cluster path), the batch threshold selection and the bitmap coder, each against its CPU oracle.
usage: python scripts/stress_pee.py [seeds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np

from codec_tcc_b200 import container, pee
from codec_tcc_b200.synth import random_payload, synth_image, synth_saturated
from oracle import bitcode_numpy as BN
from oracle import pee_c as PC
from oracle import pee_numpy as PN

seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
n_cases = 0
for seed in range(seeds):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        h, w = int(rng.integers(3, 700)), int(rng.integers(3, 900))
        bd, itemsize = [(8, 1), (8, 2), (10, 2), (12, 2), (16, 2), (5, 1)][int(rng.integers(0, 6))]
        maxval = (1 << bd) - 1
        gen = synth_saturated if rng.integers(0, 3) == 0 else synth_image
        img = gen(h, w, maxval, int(rng.integers(0, 1 << 30))).astype(np.uint8 if itemsize == 1 else np.uint16)
        T = int(rng.integers(1, min(1 << (bd - 1), 40) + 1))
        pay = random_payload(h * w, seed)
        cap = PC.embed(img, pay, h * w, T, bd)[2]["capacity"]
        n_bits = int(cap * rng.random())
        m0, lm0, i0 = PC.embed(img, pay, n_bits, T, bd)
        i0.pop("status")
        for path in ("0", "1"):
            os.environ["PEEB_CLUSTER"] = path
            m1, lm1, i1 = pee.pee_embed(img, pay, T, bd, n_bits=n_bits)
            assert i1 == i0 and np.array_equal(m1, m0) and np.array_equal(lm1, lm0), (seed, h, w, bd, T, n_bits, path)
            p1, r1 = pee.pee_extract(m1, lm1, T, n_bits, bd)
            assert np.array_equal(r1, img) and np.array_equal(np.unpackbits(p1)[:n_bits], np.unpackbits(pay)[:n_bits]), (seed, h, w, path)
            n_cases += 1
        os.environ.pop("PEEB_CLUSTER")
        # the location map through the bitmap coder
        blob = container.encode_bitmap(lm0, packed=True)
        assert blob == BN.encode(lm0, packed=True) and np.array_equal(container.decode_bitmap(blob, lm0.size * 8, packed=True), lm0.ravel())
    # batch threshold selection on small images
    n = 5
    h, w = int(rng.integers(20, 120)), int(rng.integers(20, 200))
    imgs = np.stack([synth_image(h, w, 4095, int(rng.integers(0, 1 << 30))) for _ in range(n)])
    caps = [PC.embed(imgs[u], np.zeros(h * w // 8 + 8, np.uint8), 0, 2048, 12)[2]["capacity"] for u in range(n)]
    nb = np.array([int(c * f) for c, f in zip(caps, rng.random(n))], np.int64)
    stride = (int(nb.max()) + 7) // 8 + 8
    stride += -stride % 4
    pays = rng.integers(0, 256, (n, stride), dtype=np.uint8)
    marked, lm, info = pee.pee_embed_batch(imgs, pays, nb, None, 12)
    for u in range(n):
        try:
            mo, lo, io = PN.pee_embed(imgs[u], pays[u], None, 12, n_bits=int(nb[u]))
        except ValueError:
            assert int(info[u, 7]) == pee.PEEB_E_CAPACITY
            continue
        assert int(info[u, 0]) == io["T"] and np.array_equal(marked[u], mo) and np.array_equal(lm[u], lo), (seed, u)
    # random bitmaps
    nmap = int(rng.integers(0, 300000))
    a = ((rng.random(nmap) < rng.random() ** 3) * rng.integers(1, 256, nmap)).astype(np.uint8)
    blob = container.encode_bitmap(a)
    assert blob == BN.encode(a) and np.array_equal(container.decode_bitmap(blob, nmap), (a != 0).astype(np.uint8))
print("stress ok:", n_cases, "embed/extract cases,", seeds, "threshold-selection batches and bitmaps")
if os.environ.get("PEEB_LIBRARY"):  # the bounds-checked build (python -m codec_tcc_b200.build --bounds): what its checker saw
    from codec_tcc_b200 import _cabi

    print("bounds:", _cabi.debug_bounds())
