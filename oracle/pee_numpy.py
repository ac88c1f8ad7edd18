"""TEST INFRASTRUCTURE ONLY -- numpy oracle for reversible Prediction-Error
Expansion (PEE), row a10 of SURVEY.md section 8.

*** PARITY UNPINNED ***  The mounted reference (wesleyfn/codec-tcc) advertises
PEE in README.md:3 but ships no PEE code, test or golden vector.
This file therefore follows the specification in SURVEY.md Appendix A, not a
reference source file; GPU == this oracle bit-for-bit, plus extract(embed(x))
== x, is all that can be claimed.  ``oracle/pee_ref.c`` is an independent
scalar C restatement of the same specification; the two are cross-checked in
``tests/test_pee_oracle.py``.

Nothing in the product package imports this file.

Conventions (Appendix A):
  * interior pixels only (1 <= i <= h-2, 1 <= j <= w-2); colour = (i+j)&1;
    embed order colour 0 then 1, extract order colour 1 then 0;
  * predictor p = (N+S+W+E) >> 2 on the current working image, e = x - p;
  * payload: packed bytes, most significant bit first (np.packbits default,
    same bit order as src/codec.py:239-240), ``n_bits`` valid bits;
  * location map: np.packbits(lm, axis=1), shape (h, ceil(w/8)).
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------
def payload_to_bits(payload, n_bits=None) -> np.ndarray:
    """-> uint8 vector of 0/1 with exactly n_bits entries."""
    if isinstance(payload, str):
        bits = np.frombuffer(payload.encode("ascii"), dtype=np.uint8) - ord("0")
        if bits.size and bits.max() > 1:
            raise ValueError("payload string must contain only '0'/'1'")
        return bits if n_bits is None else bits[:n_bits]
    packed = np.frombuffer(bytes(payload), dtype=np.uint8) if not isinstance(payload, np.ndarray) \
        else np.ascontiguousarray(payload, dtype=np.uint8).reshape(-1)
    bits = np.unpackbits(packed)
    if n_bits is None:
        n_bits = bits.size
    if n_bits > bits.size:
        raise ValueError("n_bits exceeds the packed payload length")
    return bits[:n_bits]


def _maxval(img, bit_depth):
    if bit_depth is None:
        bit_depth = 8 * img.dtype.itemsize
    return bit_depth, (1 << bit_depth) - 1


def _colour_mask(h, w, colour):
    """Boolean (h-2, w-2) mask of interior pixels with (i+j)&1 == colour."""
    ii = np.arange(1, h - 1)[:, None]
    jj = np.arange(1, w - 1)[None, :]
    return ((ii + jj) & 1) == colour


def _predict(cur):
    """Rhombus predictor over the interior of an int64 working image."""
    return (cur[:-2, 1:-1] + cur[2:, 1:-1] + cur[1:-1, :-2] + cur[1:-1, 2:]) >> 2


# --------------------------------------------------------------------------
def _embed_pass(cur, lm, colour, T, maxval, bits, base):
    """One colour pass in place on ``cur`` / ``lm``; returns #carriers."""
    h, w = cur.shape
    sel = _colour_mask(h, w, colour)
    x = cur[1:-1, 1:-1]
    p = _predict(cur)
    e = x - p
    expand = sel & (e >= -T) & (e < T)
    up = sel & (e >= T)
    down = sel & (e < -T)
    v = p + 2 * e
    flag = (expand & ((v < 0) | (v + 1 > maxval))) | (up & (x + T > maxval)) | (down & (x - T < 0))
    carrier = expand & ~flag
    k = np.cumsum(carrier.reshape(-1)).reshape(carrier.shape) - 1  # raster order
    ncar = int(carrier.sum())
    K = base + k[carrier]
    b = np.zeros(ncar, dtype=np.int64)
    inside = K < bits.size
    b[inside] = bits[K[inside]]
    new = x.copy()
    new[carrier] = v[carrier] + b
    mv_up = up & ~flag
    mv_dn = down & ~flag
    new[mv_up] = x[mv_up] + T
    new[mv_dn] = x[mv_dn] - T
    cur[1:-1, 1:-1] = new
    lm[1:-1, 1:-1] |= flag.astype(np.uint8)
    return ncar


def embed_fixed_T(img, bits, T, maxval):
    """Both passes at a given T.  Returns (marked int64, lm uint8, cap0, cap1)
    without checking capacity."""
    h, w = img.shape
    cur = img.astype(np.int64)
    lm = np.zeros((h, w), dtype=np.uint8)
    if h < 3 or w < 3:
        return cur, lm, 0, 0
    cap0 = _embed_pass(cur, lm, 0, T, maxval, bits, 0)
    cap1 = _embed_pass(cur, lm, 1, T, maxval, bits, cap0)
    return cur, lm, cap0, cap1


def error_histogram(img, bit_depth=None):
    """Appendix A 'threshold selection': per-colour histogram of prediction
    errors on the ORIGINAL image over interior pixels that are not flagged for
    expansion.  Returns int64 array (2, 2*Tmax) indexed by e + Tmax for
    -Tmax <= e < Tmax (errors outside that window can never be expandable and
    are dropped), Tmax = 2**(bit_depth-1)."""
    bit_depth, maxval = _maxval(img, bit_depth)
    tmax = 1 << (bit_depth - 1)
    hist = np.zeros((2, 2 * tmax), dtype=np.int64)
    h, w = img.shape
    if h < 3 or w < 3:
        return hist
    cur = img.astype(np.int64)
    x = cur[1:-1, 1:-1]
    p = _predict(cur)
    e = x - p
    v = p + 2 * e
    ok = ~((v < 0) | (v + 1 > maxval)) & (e >= -tmax) & (e < tmax)
    for c in (0, 1):
        sel = _colour_mask(h, w, c) & ok
        hist[c] = np.bincount((e[sel] + tmax).astype(np.int64), minlength=2 * tmax)
    return hist


def estimate_T(hist, n_bits):
    """min{T >= 1 : sum_c sum_{-T<=e<T} hist_c[e] >= n_bits}; None if no T up
    to Tmax qualifies."""
    tmax = hist.shape[1] // 2
    tot = hist.sum(axis=0)
    # est(T) = sum_{e=-T}^{T-1}: grow symmetric window around index tmax
    left = np.cumsum(tot[:tmax][::-1])   # e = -1, -2, ..., -Tmax
    right = np.cumsum(tot[tmax:])        # e = 0, 1, ..., Tmax-1
    est = left + right                   # est[T-1]
    ok = np.flatnonzero(est >= n_bits)
    return None if ok.size == 0 else int(ok[0]) + 1


def pee_embed(img, payload, T=None, bit_depth=None, n_bits=None):
    """-> (marked, lm_packed, info).  Raises ValueError when the payload does
    not fit (n_bits > capacity)."""
    img = np.ascontiguousarray(img)
    if img.ndim != 2 or img.dtype not in (np.uint8, np.uint16):
        raise ValueError("image must be 2-D uint8/uint16")
    bit_depth, maxval = _maxval(img, bit_depth)
    bits = payload_to_bits(payload, n_bits)
    n_bits = int(bits.size)
    tmax = 1 << (bit_depth - 1)
    if T is None:
        T = estimate_T(error_histogram(img, bit_depth), n_bits)
        if T is None:
            raise ValueError("payload exceeds capacity at every threshold")
        while True:
            cur, lm, cap0, cap1 = embed_fixed_T(img, bits, T, maxval)
            if cap0 + cap1 >= n_bits:
                break
            T += 1
            if T > tmax:
                raise ValueError("payload exceeds capacity at every threshold")
    else:
        T = int(T)
        if T < 1 or T > tmax:
            raise ValueError("T out of range")
        cur, lm, cap0, cap1 = embed_fixed_T(img, bits, T, maxval)
        if n_bits > cap0 + cap1:
            raise ValueError(f"payload of {n_bits} bits exceeds capacity {cap0 + cap1} at T={T}")
    diff = cur - img.astype(np.int64)
    info = {
        "T": T, "n_bits": n_bits, "capacity": cap0 + cap1, "cap0": cap0, "cap1": cap1,
        "n_flagged": int(lm.sum()), "sse": int((diff * diff).sum()),
    }
    return cur.astype(img.dtype), np.packbits(lm, axis=1), info


# --------------------------------------------------------------------------
def _extract_pass(cur, lm, colour, T):
    """Undo one colour pass in place; returns the carrier bits in raster order."""
    h, w = cur.shape
    sel = _colour_mask(h, w, colour) & (lm[1:-1, 1:-1] == 0)
    x = cur[1:-1, 1:-1]
    p = _predict(cur)
    ee = x - p
    carrier = sel & (ee >= -2 * T) & (ee < 2 * T)
    up = sel & (ee >= 2 * T)
    down = sel & (ee < -2 * T)
    e = ee.copy()
    e[carrier] = ee[carrier] >> 1
    e[up] = ee[up] - T
    e[down] = ee[down] + T
    out_bits = (ee[carrier] & 1).astype(np.uint8)  # boolean indexing is raster order
    new = x.copy()
    new[sel] = (p + e)[sel]
    cur[1:-1, 1:-1] = new
    return out_bits


def pee_extract(marked, lm_packed, T, n_bits, bit_depth=None):
    """-> (payload_packed uint8[ceil(n_bits/8)], recovered)."""
    marked = np.ascontiguousarray(marked)
    h, w = marked.shape
    lm = np.unpackbits(np.ascontiguousarray(lm_packed, dtype=np.uint8), axis=1)[:, :w] if w else \
        np.zeros((h, 0), np.uint8)
    cur = marked.astype(np.int64)
    if h < 3 or w < 3:
        bits1 = bits0 = np.zeros(0, np.uint8)
    else:
        bits1 = _extract_pass(cur, lm, 1, int(T))
        bits0 = _extract_pass(cur, lm, 0, int(T))
    allbits = np.concatenate([bits0, bits1])
    if n_bits > allbits.size:
        raise ValueError("n_bits exceeds the number of carriers found")
    return np.packbits(allbits[:n_bits]), cur.astype(marked.dtype)


def pee_sweep(img, payload, T_values, bit_depth=None, n_bits=None):
    """Capacity / distortion table: one real embed per T (pass 1 depends on the
    output of pass 0, so a histogram shortcut would not be exact).  The payload
    is zero-padded / truncated to the capacity at each T.
    -> list of dicts {T, capacity, cap0, cap1, n_flagged, sse, mse, psnr}."""
    img = np.ascontiguousarray(img)
    bit_depth, maxval = _maxval(img, bit_depth)
    bits = payload_to_bits(payload, n_bits)
    rows = []
    for T in T_values:
        cur, lm, cap0, cap1 = embed_fixed_T(img, bits, int(T), maxval)
        diff = cur - img.astype(np.int64)
        sse = int((diff * diff).sum())
        mse = sse / img.size if img.size else 0.0
        psnr = float("inf") if sse == 0 else 10 * np.log10((maxval ** 2) / mse)
        rows.append({"T": int(T), "capacity": cap0 + cap1, "cap0": cap0, "cap1": cap1,
                     "n_flagged": int(lm.sum()), "sse": sse, "mse": mse, "psnr": float(psnr)})
    return rows


# --------------------------------------------------------------------------
# N1: causal MED predictor, DESIGN.md "Appendix A2" -- *** PARITY UNPINNED ***.
# Second, vectorised formulation of oracle/pee_ref.c:pee_med_ref_embed (masks + cumsum); the embedder
# predicts from the ORIGINAL pixels, so it needs no sequential walk.  Extraction is inherently
# sequential (it predicts from recovered pixels): only the C oracle restates it.
# --------------------------------------------------------------------------
def med_predict(img64):
    """MED (JPEG-LS) prediction for pixels (1.., 1..) of an int64 image: a = W, b = N, c = NW."""
    a, b, c = img64[1:, :-1], img64[:-1, 1:], img64[:-1, :-1]
    return np.clip(a + b - c, np.minimum(a, b), np.maximum(a, b))


def med_embed(img, payload, T, bit_depth=None, n_bits=None):
    """-> (marked, lm_packed, info) like the C oracle (no exception on overflow: info['status'] = -2)."""
    img = np.ascontiguousarray(img)
    h, w = img.shape
    _, maxval = _maxval(img, bit_depth)
    bits = payload_to_bits(payload, n_bits)
    n_bits = bits.size
    cur = img.astype(np.int64)
    lm = np.zeros((h, w), np.uint8)
    cap = 0
    if h >= 2 and w >= 2:
        x = cur[1:, 1:]
        p = med_predict(cur)
        e = x - p
        expd = (e >= -T) & (e < T)
        v = p + 2 * e
        flag = np.where(expd, (v < 0) | (v + 1 > maxval), np.where(e >= T, x + T > maxval, x - T < 0))
        carrier = expd & ~flag
        rank = np.cumsum(carrier.reshape(-1)) - 1          # raster order
        cap = int(carrier.sum())
        padded = np.zeros(max(cap, 1), np.int64)
        padded[:min(cap, n_bits)] = bits[:min(cap, n_bits)]
        b = np.where(carrier.reshape(-1), padded[np.clip(rank, 0, None)], 0).reshape(x.shape)
        y = np.where(flag, x, np.where(expd, v + b, np.where(e >= T, x + T, x - T)))
        out = cur.copy()
        out[1:, 1:] = y
        lm[1:, 1:] = flag
        cur_out = out
    else:
        cur_out = cur
    marked = cur_out.astype(img.dtype)
    sse = int(((cur_out - cur) ** 2).sum())
    info = {"T": int(T), "n_bits": int(n_bits), "capacity": cap, "cap0": cap, "cap1": 0, "n_flagged": int(lm.sum()),
            "sse": sse, "status": -2 if n_bits > cap else 0}
    return marked, np.packbits(lm, axis=1), info
