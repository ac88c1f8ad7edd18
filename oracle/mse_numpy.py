"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64) of the
reference's distortion metrics, rows a1-a4 of SURVEY.md section 8
(``AnalisadorMSE`` in src/mse.py), for *array* inputs.

The product computes these from exact integer moments on the GPU; this file
keeps the reference's own float64 element-wise arithmetic so that the two
independent formulations can be compared.  Pinned against the unmodified
reference by ``tests/test_oracle_vs_reference.py`` and ``tests/golden``.
"""
from __future__ import annotations

import numpy as np


def _as_float_with_range(img):
    """src/mse.py:85-87 -- arrays are converted to float64 and their *own
    maximum* is taken as the value range."""
    f = np.array(img, dtype=np.float64)
    return f, f.max()


def _normalise(f1, r1, f2, r2):
    """src/mse.py:101-110 / :153-161 -- when the two maxima differ both images
    are rescaled to the larger one in floating point."""
    if r1 != r2:
        top = max(r1, r2)
        return (f1 / r1) * top, (f2 / r2) * top, top
    return f1, f2, max(r1, r2)


def calcular_mse(img1, img2):
    """src/mse.py:74-116 -> ``(mse, max_range)``."""
    f1, r1 = _as_float_with_range(img1)
    f2, r2 = _as_float_with_range(img2)
    if f1.shape != f2.shape:
        raise ValueError(f"Dimensões diferentes: {f1.shape} vs {f2.shape}")
    g1, g2, top = _normalise(f1, r1, f2, r2)
    if r1 == r2:
        top = r1  # :110
    return np.mean((g1 - g2) ** 2), top


def calcular_psnr(mse, max_valor=None):
    """src/mse.py:118-133."""
    if mse == 0:
        return float("inf")
    peak = 255 if max_valor is None else max_valor
    return 10 * np.log10((peak ** 2) / mse)


def calcular_ssim_simples(img1, img2):
    """src/mse.py:135-179 -- one global window."""
    f1, r1 = _as_float_with_range(img1)
    f2, r2 = _as_float_with_range(img2)
    g1, g2, top = _normalise(f1, r1, f2, r2)
    m1, m2 = np.mean(g1), np.mean(g2)
    v1, v2 = np.var(g1), np.var(g2)
    cov = np.mean((g1 - m1) * (g2 - m2))
    k1 = (0.01 * top) ** 2
    k2 = (0.03 * top) ** 2
    return ((2 * m1 * m2 + k1) * (2 * cov + k2)) / ((m1 ** 2 + m2 ** 2 + k1) * (v1 + v2 + k2))


def difference_stats(img1, img2):
    """src/mse.py:202-209 -> (mean |d|, max |d|, #changed, percent changed)."""
    a = np.array(img1, dtype=np.float64)
    b = np.array(img2, dtype=np.float64)
    ad = np.abs(a - b)
    changed = np.sum(a != b)
    return np.mean(ad), np.max(ad), changed, (changed / a.size) * 100
