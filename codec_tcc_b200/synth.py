"""Seeded synthetic DICOM-shaped images: a smooth field
plus Gaussian noise, clipped to the stored bit depth.  Used by the tests, the
golden-vector generator and bench.py; there is no network for real datasets."""
from __future__ import annotations

import numpy as np


def synth_image(h: int, w: int, maxval: int, seed: int) -> np.ndarray:
    """((sin(x/97)+cos(y/61))*0.25+0.5)*maxval*0.6 + N(0, maxval/256), clipped
    to [0, maxval]; uint8 when maxval <= 255 else uint16."""
    rng = np.random.default_rng(seed)
    y = np.arange(h, dtype=np.float64)[:, None]
    x = np.arange(w, dtype=np.float64)[None, :]
    field = ((np.sin(x / 97.0) + np.cos(y / 61.0)) * 0.25 + 0.5) * maxval * 0.6
    noisy = field + rng.normal(0.0, maxval / 256.0, size=(h, w))
    dtype = np.uint8 if maxval <= 255 else np.uint16
    return np.clip(np.rint(noisy), 0, maxval).astype(dtype)


def synth_batch(n: int, h: int, w: int, maxval: int, seed: int) -> np.ndarray:
    """n independent images, image k seeded with ``seed + k`` -> (n, h, w)."""
    dtype = np.uint8 if maxval <= 255 else np.uint16
    out = np.empty((n, h, w), dtype=dtype)
    for k in range(n):
        out[k] = synth_image(h, w, maxval, seed + k)
    return out


def synth_saturated(h: int, w: int, maxval: int, seed: int) -> np.ndarray:
    """An image with large clipped (0 / maxval) regions and hard edges: the
    stress case for the PEE overflow/underflow location map."""
    rng = np.random.default_rng(seed)
    img = synth_image(h, w, maxval, seed).astype(np.int64)
    img = np.rint((img - maxval * 0.3) * 2.5).astype(np.int64)
    blocks = rng.integers(0, 3, size=((h + 15) // 16, (w + 15) // 16))
    sel = np.kron(blocks, np.ones((16, 16), dtype=np.int64))[:h, :w]
    img = np.where(sel == 0, img, np.where(sel == 1, maxval - (img % 3), img % 3))
    dtype = np.uint8 if maxval <= 255 else np.uint16
    return np.clip(img, 0, maxval).astype(dtype)


def random_payload(n_bits: int, seed: int) -> np.ndarray:
    """Packed (MSB-first) random payload of n_bits bits; pad bits are zero."""
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, size=n_bits, dtype=np.uint8)
    return np.packbits(bits)
