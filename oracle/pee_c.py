"""ctypes front end of oracle/libpee_oracle.so (the scalar C PEE oracle).
TEST INFRASTRUCTURE ONLY; *** PARITY UNPINNED *** (see pee_ref.c)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_oracle

_lib = None

INFO_KEYS = ("T", "n_bits", "capacity", "cap0", "cap1", "n_flagged", "sse", "status")


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.OUT
        if not os.path.exists(path):
            path = build_oracle.build()
        L = C.CDLL(path)
        vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
        L.pee_ref_embed.argtypes = [vp, i32, i32, i32, i64, i32, vp, i64, vp, vp, vp]
        L.pee_ref_embed.restype = i32
        L.pee_ref_extract.argtypes = [vp, i32, i32, i32, i32, vp, i64, vp, vp]
        L.pee_ref_extract.restype = i32
        L.pee_ref_hist.argtypes = [vp, i32, i32, i32, i64, i64, vp]
        L.pee_ref_hist.restype = None
        L.pee_ref_embed_batch.argtypes = [vp, i32, i32, i32, i32, i64, i32, vp, i64, vp, vp, vp, vp, i32]
        L.pee_ref_embed_batch.restype = i32
        L.pee_ref_extract_batch.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, i64, vp, i32]
        L.pee_ref_extract_batch.restype = i32
        L.pee_ref_threads.restype = i32
        L.pee_med_ref_embed.argtypes = [vp, i32, i32, i32, i64, i32, vp, i64, vp, vp, vp]
        L.pee_med_ref_embed.restype = i32
        L.pee_med_ref_extract.argtypes = [vp, i32, i32, i32, i32, vp, i64, vp, vp]
        L.pee_med_ref_extract.restype = i32
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _maxval(img, bit_depth):
    bd = 8 * img.dtype.itemsize if bit_depth is None else bit_depth
    return bd, (1 << bd) - 1


def embed(img, payload_packed, n_bits, T, bit_depth=None, predictor="rhombus"):
    """-> (marked, lm_packed, info dict incl. 'status').  Does not raise on
    overflow; status == -2 then.  predictor: "rhombus" (Appendix A) or "med" (causal, N1)."""
    img = np.ascontiguousarray(img)
    h, w = img.shape
    _, maxval = _maxval(img, bit_depth)
    pay = np.ascontiguousarray(payload_packed, dtype=np.uint8).reshape(-1)
    if pay.size * 8 < n_bits:
        raise ValueError("n_bits exceeds packed payload")
    if pay.size == 0:
        pay = np.zeros(1, np.uint8)
    marked = np.empty_like(img)
    lm = np.empty((h, (w + 7) // 8), np.uint8)
    info = np.zeros(8, np.int64)
    fn = lib().pee_med_ref_embed if predictor == "med" else lib().pee_ref_embed
    fn(_p(img), h, w, img.dtype.itemsize, maxval, int(T), _p(pay), int(n_bits), _p(marked), _p(lm), _p(info))
    return marked, lm, dict(zip(INFO_KEYS, (int(v) for v in info)))


def extract(marked, lm_packed, T, n_bits, predictor="rhombus"):
    marked = np.ascontiguousarray(marked)
    h, w = marked.shape
    lm = np.ascontiguousarray(lm_packed, dtype=np.uint8)
    out = np.zeros(max(1, (n_bits + 7) // 8), np.uint8)
    rec = np.empty_like(marked)
    fn = lib().pee_med_ref_extract if predictor == "med" else lib().pee_ref_extract
    rc = fn(_p(marked), h, w, marked.dtype.itemsize, int(T), _p(lm), int(n_bits), _p(out), _p(rec))
    if rc != 0:
        raise ValueError("n_bits exceeds the number of carriers found")
    return out[: (n_bits + 7) // 8], rec


def hist(img, bit_depth=None):
    img = np.ascontiguousarray(img)
    h, w = img.shape
    bd, maxval = _maxval(img, bit_depth)
    tmax = 1 << (bd - 1)
    out = np.zeros((2, 2 * tmax), np.int64)
    lib().pee_ref_hist(_p(img), h, w, img.dtype.itemsize, maxval, tmax, _p(out))
    return out


def embed_batch(imgs, payloads, n_bits, T, bit_depth=None, threads=0):
    """imgs (n,h,w); payloads (n, stride) uint8; n_bits (n,) int64."""
    imgs = np.ascontiguousarray(imgs)
    n, h, w = imgs.shape
    _, maxval = _maxval(imgs, bit_depth)
    payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
    nb = np.ascontiguousarray(n_bits, dtype=np.int64)
    marked = np.empty_like(imgs)
    lm = np.empty((n, h, (w + 7) // 8), np.uint8)
    info = np.zeros((n, 8), np.int64)
    lib().pee_ref_embed_batch(_p(imgs), n, h, w, imgs.dtype.itemsize, maxval, int(T), _p(payloads),
                              payloads.shape[1], _p(nb), _p(marked), _p(lm), _p(info), threads)
    return marked, lm, info


def extract_batch(marked, lm, T, n_bits, payload_stride, threads=0):
    marked = np.ascontiguousarray(marked)
    n, h, w = marked.shape
    nb = np.ascontiguousarray(n_bits, dtype=np.int64)
    out = np.zeros((n, payload_stride), np.uint8)
    rec = np.empty_like(marked)
    rc = lib().pee_ref_extract_batch(_p(marked), n, h, w, marked.dtype.itemsize, int(T), _p(lm), _p(nb),
                                     _p(out), payload_stride, _p(rec), threads)
    return out, rec, rc


def threads():
    return lib().pee_ref_threads()
