"""Distortion metrics of the reference's ``src/mse.py`` (class ``AnalisadorMSE``),
rows a1-a4, computed from exact integer moments reduced on the GPU.

One streaming kernel pass over the two images yields sum (a-b)^2, sum |a-b|,
max |a-b|, #changed, sum a, sum b, sum a^2, sum b^2, sum ab, max a, max b as
int64.  From those:
  * MSE = SSE / N is *exactly* the reference's float64 ``np.mean((a-b)**2)``
    (every partial sum of integer squares is an exact float64 below 2^53);
  * PSNR is the reference's scalar formula;
  * the range-normalisation branch (src/mse.py:101-106) and the global SSIM
    (src/mse.py:164-178) are evaluated with exact rational arithmetic on the
    moments, which agrees with the reference's float64 element-wise evaluation
    to ~1e-13 relative (tests allow 1e-9).
Only array inputs of unsigned 8/16-bit pixels are handled on the device; file
paths need the reference's pydicom/Pillow loader, which is host I/O and out of
scope.
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

from . import _cabi
from . import dicom_raw
from ._cabi import MOMENTS, check, lib, ptr, workspace

__all__ = ["AnalisadorMSE", "image_moments"]

_KEYS = ("sse", "sad", "max_abs", "changed", "sum_a", "sum_b", "sum_aa", "sum_bb", "sum_ab", "max_a", "max_b", "n")


def _as_pixels(img, name):
    """uint8/uint16 C-contiguous array with the values of ``img``; mirrors the
    reference's ``np.array(img, dtype=np.float64)`` acceptance of lists etc.
    (src/mse.py:85) for the integer inputs the device path supports."""
    if isinstance(img, str):
        raise NotImplementedError(
            "only uncompressed .dcm paths are read here (codec_tcc_b200.dicom_raw); other files go through the "
            "reference's pydicom/PIL loader (src/mse.py:13-72), which is host I/O outside this package: load the "
            "pixels and pass the array")
    a = np.asarray(img)
    if a.dtype in (np.uint8, np.uint16):
        return np.ascontiguousarray(a)
    if a.dtype == np.bool_:
        return np.ascontiguousarray(a.astype(np.uint8))
    if np.issubdtype(a.dtype, np.integer):
        if a.size and (a.min() < 0 or a.max() > 65535):
            return None  # wider than 16 bits or negative: the float64 kernels
        return np.ascontiguousarray(a.astype(np.uint16))
    if np.issubdtype(a.dtype, np.floating):
        r = np.rint(a)
        if a.size and (not np.array_equal(r, a) or a.min() < 0 or a.max() > 65535):
            return None  # fractional / negative / wide values: the float64 kernels
        return np.ascontiguousarray(r.astype(np.uint16))
    raise TypeError(f"{name}: unsupported dtype {a.dtype}")


def _float_moments(img1, img2, device=None) -> dict:
    """The float64 route of the reference (``np.array(img, dtype=np.float64)``, src/mse.py:85,91) for pixel data
    that is not integer-valued 8/16-bit: plain sums and difference statistics, plus the normalised, centred second
    moments (src/mse.py:100-110, :152-168), all reduced on the device."""
    a = np.ascontiguousarray(np.array(img1, dtype=np.float64))
    b = np.ascontiguousarray(np.array(img2, dtype=np.float64))
    if a.shape != b.shape:
        raise ValueError(f"Dimensões diferentes: {a.shape} vs {b.shape}")  # src/mse.py:97-98
    if a.size == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    out = np.zeros(26, np.float64)
    check(lib().peeb_moments_f64_h(workspace(device).handle, ptr(a), ptr(b), a.size, ptr(out)), "peeb_moments_f64_h")
    n = a.size
    p1, p2, mu0, mv0 = out[:12], out[12:24], out[24], out[25]
    mu, mv = p2[0] / n, p2[1] / n                       # exact means of the (normalised) images
    return {
        "float": True, "n": n, "max_a": p1[8], "max_b": p1[9], "sad": p1[6], "max_abs": p1[7], "changed": int(p1[10]),
        "mu1": mu, "mu2": mv,
        "var1": p2[2] / n - (mu - mu0) ** 2, "var2": p2[3] / n - (mv - mv0) ** 2,
        "cov": p2[4] / n - (mu - mu0) * (mv - mv0), "mse": p2[5] / n,
    }


def image_moments(img1, img2, device=None, full=True) -> dict:
    """The twelve integers (see module docstring) as Python ints.  ``full=False``
    runs the lighter SSE-only kernel: only sse, max_a, max_b and n are filled."""
    # file paths: the reference takes the value range from the file (2**BitsStored - 1) instead of the
    # array maximum (src/mse.py:82-84 vs :85-87); uncompressed .dcm files are read without pydicom
    ra = rb = None
    if isinstance(img1, str) and img1.lower().endswith(".dcm"):
        img1, info = dicom_raw.read_pixels(img1)
        ra = (1 << info["BitsStored"]) - 1
    if isinstance(img2, str) and img2.lower().endswith(".dcm"):
        img2, info = dicom_raw.read_pixels(img2)
        rb = (1 << info["BitsStored"]) - 1
    a = _as_pixels(img1, "img1")
    b = _as_pixels(img2, "img2")
    if a is None or b is None:
        if ra is not None or rb is not None:
            raise TypeError("a .dcm path can only be compared with integer pixel data")
        return _float_moments(img1, img2, device)
    if a.shape != b.shape:
        raise ValueError(f"Dimensões diferentes: {a.shape} vs {b.shape}")  # src/mse.py:97-98
    if a.size == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    if a.dtype != b.dtype:  # one uint8, one uint16: widen
        a = a.astype(np.uint16)
        b = b.astype(np.uint16)
    out = np.zeros(MOMENTS, np.int64)
    ws = workspace(device)
    fn = lib().peeb_moments_h if full else lib().peeb_sse_h
    check(fn(ws.handle, ptr(a), ptr(b), a.size, a.dtype.itemsize, ptr(out)), "peeb_moments_h")
    m = {k: int(v) for k, v in zip(_KEYS, out)}
    if ra is not None:
        m["max_a"] = ra
    if rb is not None:
        m["max_b"] = rb
    return m


def _scales(m):
    """Range handling shared by calcular_mse and calcular_ssim_simples
    (src/mse.py:100-110, :152-161): returns (alpha, beta, max_range) with
    img1_norm = alpha*img1 and img2_norm = beta*img2 as exact rationals."""
    r1, r2 = m["max_a"], m["max_b"]
    if r1 == r2:
        return Fraction(1), Fraction(1), r1
    top = max(r1, r2)
    # a zero maximum makes the reference divide by zero (nan/inf); mirror with an error
    if r1 == 0 or r2 == 0:
        raise ZeroDivisionError("one image is all zeros: the reference's range normalisation divides by zero")
    return Fraction(top, r1), Fraction(top, r2), top


class AnalisadorMSE:
    """Drop-in for the numeric methods of the reference's ``AnalisadorMSE``
    (src/mse.py:9) on array inputs."""

    def __init__(self, device=None):
        self.resultados = []
        self._device = device

    def carregar_imagem(self, caminho):
        """src/mse.py:13-37 for uncompressed .dcm files: ``(float64 array, max_valor, bits_stored)``."""
        if not caminho.lower().endswith(".dcm"):
            raise NotImplementedError("only uncompressed .dcm files are read without the reference's PIL loader (src/mse.py:39-72)")
        return dicom_raw.carregar_imagem(caminho)

    # -- a1 ---------------------------------------------------------------
    def calcular_mse(self, imagem1, imagem2):
        """src/mse.py:74-116 -> ``(np.float64 mse, np.float64 max_range)``."""
        m = image_moments(imagem1, imagem2, self._device, full=False)
        if not m.get("float") and m["max_a"] != m["max_b"]:  # the normalisation branch needs the second moments too
            m = image_moments(imagem1, imagem2, self._device)
        return self._mse_from(m)

    @staticmethod
    def _mse_from(m):
        if m.get("float"):
            return np.float64(m["mse"]), np.float64(max(m["max_a"], m["max_b"]))
        al, be, top = _scales(m)
        n = m["n"]
        if al == 1 and be == 1:
            mse = np.float64(m["sse"]) / np.float64(n)  # == np.mean of exact integer squares
        else:
            # sum (al*a - be*b)^2 = al^2 Saa - 2 al be Sab + be^2 Sbb, exactly
            tot = al * al * m["sum_aa"] - 2 * al * be * m["sum_ab"] + be * be * m["sum_bb"]
            mse = np.float64(float(tot / n))
        return mse, np.float64(top)

    # -- a2 ---------------------------------------------------------------
    def calcular_psnr(self, mse, max_valor=None):
        """src/mse.py:118-133 (host scalar)."""
        if mse == 0:
            return float("inf")
        if max_valor is None:
            max_valor = 255
        return 10 * np.log10((max_valor ** 2) / mse)

    # -- a3 ---------------------------------------------------------------
    def calcular_ssim_simples(self, imagem1, imagem2):
        """src/mse.py:135-179: single-window SSIM from the five moments."""
        m = image_moments(imagem1, imagem2, self._device)
        return self._ssim_from(m)

    @staticmethod
    def _ssim_from(m):
        if m.get("float"):  # src/mse.py:152-178 on the device-reduced float64 moments
            top_f = float(max(m["max_a"], m["max_b"]))
            c1, c2 = (0.01 * top_f) ** 2, (0.03 * top_f) ** 2
            num = (2 * m["mu1"] * m["mu2"] + c1) * (2 * m["cov"] + c2)
            den = (m["mu1"] ** 2 + m["mu2"] ** 2 + c1) * (m["var1"] + m["var2"] + c2)
            return np.float64(num / den)
        al, be, top = _scales(m)
        n = m["n"]
        mu1 = al * Fraction(m["sum_a"], n)
        mu2 = be * Fraction(m["sum_b"], n)
        var1 = al * al * Fraction(n * m["sum_aa"] - m["sum_a"] ** 2, n * n)
        var2 = be * be * Fraction(n * m["sum_bb"] - m["sum_b"] ** 2, n * n)
        cov = al * be * Fraction(n * m["sum_ab"] - m["sum_a"] * m["sum_b"], n * n)
        f = float
        top_f = float(top)
        c1 = (0.01 * top_f) ** 2
        c2 = (0.03 * top_f) ** 2
        num = (2 * f(mu1) * f(mu2) + c1) * (2 * f(cov) + c2)
        den = (f(mu1) ** 2 + f(mu2) ** 2 + c1) * (f(var1) + f(var2) + c2)
        return np.float64(num / den)

    # -- a4 ---------------------------------------------------------------
    def difference_stats(self, imagem1, imagem2):
        """The difference statistics of src/mse.py:202-209 ->
        (mean |d|, max |d|, #changed, percent changed)."""
        m = image_moments(imagem1, imagem2, self._device)
        n = m["n"]
        return (np.float64(m["sad"]) / np.float64(n), np.float64(m["max_abs"]), np.int64(m["changed"]),
                (np.int64(m["changed"]) / n) * 100)

    def analisar_par_arrays(self, original, stego, nome_par=""):
        """The result record of ``analisar_par_imagens`` (src/mse.py:244-254) for
        two arrays, from ONE pass over the data (the reference makes three
        passes and reloads the files each time, src/mse.py:190-199)."""
        m = image_moments(original, stego, self._device)
        mse, top = self._mse_from(m)
        n = m["n"]
        res = {
            "nome": nome_par, "original": None, "stego": None,
            "mse": mse, "psnr": self.calcular_psnr(mse, top), "ssim": self._ssim_from(m),
            "diferenca_media": np.float64(m["sad"]) / np.float64(n),
            "diferenca_max": np.float64(m["max_abs"]),
            "percentual_mudanca": (np.int64(m["changed"]) / n) * 100,
        }
        self.resultados.append(res)
        return res
