"""Reversible Prediction-Error Expansion on the GPU (row a10).

The reference repository advertises PEE (README.md:3) but contains no PEE
code, so these entry points follow SURVEY.md Appendix A; the
API keeps the reference's conventions: numpy uint8/uint16 2-D arrays in and
out, inputs never mutated, ``ValueError`` on bad input, payload bits most
significant first (src/codec.py:239-240).

Every function here runs hand-written sm_100a kernels through the C ABI of
``include/peeb200.h``; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from ._cabi import INFO, INFO_KEYS, PEEB_E_CAPACITY, check, lib, ptr, workspace

__all__ = [
    "pee_embed", "pee_extract", "pee_sweep", "pee_sweep_pairs", "pee_histogram", "pee_embed_batch",
    "pee_extract_batch", "pack_payload", "estimate_threshold",
]


# ---------------------------------------------------------------- payloads
def pack_payload(payload, n_bits=None):
    """-> (packed uint8 array, n_bits).  Accepts a '0'/'1' string (the
    reference's ``message_to_bits`` format), bytes, or a uint8 array of packed
    bits (MSB first)."""
    if isinstance(payload, str):
        raw = np.frombuffer(payload.encode("ascii"), dtype=np.uint8)
        bits = raw - ord("0")
        if bits.size and bits.max() > 1:
            raise ValueError("payload string must contain only '0'/'1'")
        if n_bits is not None:
            if n_bits > bits.size:
                raise ValueError("n_bits exceeds the payload length")
            bits = bits[:n_bits]
        return np.packbits(bits), int(bits.size)
    packed = np.frombuffer(bytes(payload), dtype=np.uint8) if not isinstance(payload, np.ndarray) \
        else np.ascontiguousarray(payload, dtype=np.uint8).reshape(-1)
    if n_bits is None:
        n_bits = 8 * packed.size
    if n_bits < 0 or n_bits > 8 * packed.size:
        raise ValueError("n_bits exceeds the packed payload length")
    return packed, int(n_bits)


def _bit_depth(img, bit_depth):
    bd = 8 * img.dtype.itemsize if bit_depth is None else int(bit_depth)
    if bd < 1 or bd > 8 * img.dtype.itemsize:
        raise ValueError(f"bit_depth {bd} does not fit {img.dtype}")
    return bd


def _check_T(T, bd):
    T = int(T)
    if T < 1 or T > (1 << (bd - 1)):
        raise ValueError(f"T={T} outside 1..{1 << (bd - 1)}")
    return T


def _info_dict(row):
    d = {k: int(v) for k, v in zip(INFO_KEYS, row)}
    return d


# ---------------------------------------------------------------- batches (numpy in / numpy out)
PREDICTORS = ("rhombus", "med")


def _check_out(arr, name, shape, dtype):
    """A caller-supplied output array goes to the C ABI as a raw pointer: it must be exactly what the copy back
    writes (shape, dtype, C order, writable), or the call is refused."""
    if not isinstance(arr, np.ndarray):
        raise ValueError(f"{name} must be a numpy array")
    if arr.dtype != np.dtype(dtype) or arr.shape != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)} and dtype {np.dtype(dtype)}, got {arr.shape} {arr.dtype}")
    if not arr.flags.c_contiguous or not arr.flags.writeable:
        raise ValueError(f"{name} must be C-contiguous and writable")
    return arr


def _check_predictor(predictor):
    if predictor not in PREDICTORS:
        raise ValueError(f"predictor must be one of {PREDICTORS}")
    return predictor


def pee_embed_batch(imgs, payloads, n_bits, T, bit_depth=None, *, shared_cover=False, shared_payload=False,
                    want_marked=True, want_lm=True, out_marked=None, out_lm=None, device=None, predictor="rhombus"):
    """Embed into a batch of equally shaped images.

    imgs      (n, h, w) uint8/uint16 -- or (h, w) with ``shared_cover=True``,
              every unit then embeds into the same cover (threshold sweep)
    payloads  (n, stride) uint8, packed MSB first; row u holds unit u's bits
              (one row of (1, stride) or (stride,) with ``shared_payload=True``)
    n_bits    (n,) ints; T scalar or (n,) ints, or None: every unit gets the smallest threshold that holds its
              payload (rhombus: histogram estimate, then verify-and-increment; med: from T = 1 upwards -- all on the
              device); see info[:, 0]
    -> (marked (n,h,w) or None, lm_packed (n,h,ceil(w/8)) or None, info (n, 8) int64)
    ``info[:, 7]`` is 0 or PEEB_E_CAPACITY (-2): nothing is raised here, the
    embed of an oversize payload is the zero-padded embed of what fits.
    ``predictor``: "rhombus" (two-pass checkerboard) or "med" (causal MED predictor,
    one raster pass; no shared cover / payload).
    """
    _check_predictor(predictor)
    imgs = _cabi.as_image(imgs, "imgs")
    if shared_cover:
        if imgs.ndim != 2:
            raise ValueError("shared_cover expects one (h, w) image")
        h, w = imgs.shape
        n = len(n_bits)
    else:
        if imgs.ndim != 3:
            raise ValueError("imgs must be (n, h, w)")
        n, h, w = imgs.shape
    bd = _bit_depth(imgs, bit_depth)
    nb = np.ascontiguousarray(n_bits, dtype=np.int64).reshape(-1)
    if nb.size != n:
        raise ValueError("n_bits must have one entry per image")
    Ts = None  # T=None: chosen per unit on the device (Appendix A threshold selection), reported in info[:, 0]
    if T is not None:
        Ts = np.ascontiguousarray(np.broadcast_to(np.asarray(T, dtype=np.int32), (n,)))
        for t in np.unique(Ts):
            _check_T(t, bd)
    elif shared_cover:
        raise ValueError("T=None needs one cover per unit")
    payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
    if shared_payload:
        payloads = payloads.reshape(1, -1)
    elif payloads.ndim != 2 or payloads.shape[0] != n:
        raise ValueError("payloads must be (n, stride)")
    if n and int(((nb + 7) // 8).max()) > payloads.shape[1]:
        raise ValueError("a payload row is shorter than its n_bits")
    if n == 0:
        return (np.empty((0, h, w), imgs.dtype), np.empty((0, h, (w + 7) // 8), np.uint8), np.empty((0, INFO), np.int64))
    marked = None
    if want_marked:
        marked = (_check_out(out_marked, "out_marked", (n, h, w), imgs.dtype) if out_marked is not None
                  else _cabi.out_empty((n, h, w), imgs.dtype))
    lm = None
    if want_lm:
        lm = (_check_out(out_lm, "out_lm", (n, h, (w + 7) // 8), np.uint8) if out_lm is not None
              else _cabi.out_empty((n, h, (w + 7) // 8), np.uint8))
    info = np.zeros((n, INFO), np.int64)
    ws = workspace(device)
    flags = (1 if shared_cover else 0) | (2 if shared_payload else 0)
    if predictor == "med":
        if flags:
            raise ValueError("shared_cover / shared_payload are not supported with predictor='med'")
        check(lib().peeb_pee_med_embed_h(ws.handle, ptr(imgs), n, h, w, imgs.dtype.itemsize, bd, ptr(Ts), ptr(nb),
                                         ptr(payloads) if payloads.size else None, payloads.shape[1], ptr(marked), ptr(lm),
                                         ptr(info)), "peeb_pee_med_embed_h")
        return marked, lm, info
    check(lib().peeb_pee_embed_h(ws.handle, ptr(imgs), flags, n, h, w, imgs.dtype.itemsize, bd,
                                 ptr(Ts) if Ts is not None else None, ptr(nb), ptr(payloads) if payloads.size else None, payloads.shape[1],
                                 ptr(marked), ptr(lm), ptr(info)), "peeb_pee_embed_h")
    return marked, lm, info


def pee_extract_batch(marked, lm, T, n_bits, bit_depth=None, *, want_recovered=True, out_recovered=None,
                      out_payload=None, device=None, predictor="rhombus"):
    """-> (payloads (n, stride) uint8, recovered (n,h,w) or None, info (n,8)).
    ``info[:, 2]`` is the number of carriers found; ``info[:, 7]`` is
    PEEB_E_CAPACITY when n_bits exceeds it (the payload row is then what could
    be read, zero padded)."""
    _check_predictor(predictor)
    marked = _cabi.as_image(marked, "marked")
    if marked.ndim != 3:
        raise ValueError("marked must be (n, h, w)")
    n, h, w = marked.shape
    bd = _bit_depth(marked, bit_depth)
    lm = np.ascontiguousarray(lm, dtype=np.uint8)
    if lm.shape != (n, h, (w + 7) // 8):
        raise ValueError(f"lm must have shape {(n, h, (w + 7) // 8)}, got {lm.shape}")
    nb = np.ascontiguousarray(n_bits, dtype=np.int64).reshape(-1)
    if nb.size != n:
        raise ValueError("n_bits must have one entry per image")
    Ts = np.ascontiguousarray(np.broadcast_to(np.asarray(T, dtype=np.int32), (n,)))
    for t in np.unique(Ts):
        _check_T(t, bd)
    stride = int(((nb + 7) // 8).max()) if n else 0
    if out_payload is not None:
        if not isinstance(out_payload, np.ndarray) or out_payload.ndim != 2 or out_payload.shape[0] != n \
                or out_payload.shape[1] < stride:
            raise ValueError(f"out_payload must be a ({n}, >= {stride}) uint8 array")
        payload = _check_out(out_payload, "out_payload", out_payload.shape, np.uint8)
        stride = payload.shape[1]
    else:
        payload = np.zeros((n, stride), np.uint8)
    if n == 0:
        return payload, np.empty((0, h, w), marked.dtype), np.empty((0, INFO), np.int64)
    rec = None
    if want_recovered:
        rec = (_check_out(out_recovered, "out_recovered", (n, h, w), marked.dtype) if out_recovered is not None
               else _cabi.out_empty((n, h, w), marked.dtype))
    info = np.zeros((n, INFO), np.int64)
    ws = workspace(device)
    # a zero-width payload array still needs a valid pointer
    pay_ptr = ptr(payload) if payload.size else ptr(np.zeros(4, np.uint8))
    fn = lib().peeb_pee_med_extract_h if predictor == "med" else lib().peeb_pee_extract_h
    check(fn(ws.handle, ptr(marked), n, h, w, marked.dtype.itemsize, bd, ptr(Ts), ptr(nb),
             ptr(lm), pay_ptr, stride, ptr(rec), ptr(info)), "peeb_pee_extract_h")
    return payload, rec, info


def pee_histogram(img, bit_depth=None, device=None) -> np.ndarray:
    """Per-colour histogram of prediction errors of the ORIGINAL image over
    interior pixels not flagged for expansion (Appendix A, threshold
    selection): int64 (2, 2*Tmax), index e + Tmax."""
    img = _cabi.as_image(img)
    if img.ndim != 2:
        raise ValueError("image must be 2-D")
    bd = _bit_depth(img, bit_depth)
    tmax = 1 << (bd - 1)
    h, w = img.shape
    out = np.zeros((2, 2 * tmax), np.uint32)
    ws = workspace(device)
    check(lib().peeb_pee_hist_h(ws.handle, ptr(img), 1, h, w, img.dtype.itemsize, bd, ptr(out)), "peeb_pee_hist_h")
    return out.astype(np.int64)


def estimate_threshold(hist, n_bits):
    """min{T >= 1 : sum_c sum_{-T <= e < T} hist_c[e] >= n_bits}, None if no T
    up to Tmax qualifies (Appendix A)."""
    tmax = hist.shape[1] // 2
    tot = hist.sum(axis=0)
    est = np.cumsum(tot[:tmax][::-1]) + np.cumsum(tot[tmax:])
    ok = np.flatnonzero(est >= n_bits)
    return None if ok.size == 0 else int(ok[0]) + 1


# ---------------------------------------------------------------- single image
def pee_embed(img, payload, T=None, bit_depth=None, n_bits=None, device=None, predictor="rhombus"):
    """-> (marked, lm_packed, info dict).  ``T=None`` picks the smallest
    threshold whose capacity holds the payload (histogram estimate, then
    verify-and-increment).  ``ValueError`` when the payload does not fit."""
    img = _cabi.as_image(img)
    if img.ndim != 2:
        raise ValueError("image must be 2-D (grayscale)")
    bd = _bit_depth(img, bit_depth)
    packed, n_bits = pack_payload(payload, n_bits)
    tmax = 1 << (bd - 1)
    pay2d = packed.reshape(1, -1)

    def run(t):
        marked, lm, info = pee_embed_batch(img[None], pay2d, [n_bits], t, bd, device=device, predictor=predictor)
        return marked[0], lm[0], info[0]

    if T is None:
        # the threshold search runs on the device: histogram estimate, then verify and increment (rhombus predictor);
        # from T = 1 upwards (causal predictor, which has no estimate) -- only units that fall short are embedded again
        marked, lm, info = pee_embed_batch(img[None], pay2d, [n_bits], None, bd, device=device, predictor=predictor)
        marked, lm, info = marked[0], lm[0], info[0]
        if info[7] == PEEB_E_CAPACITY:
            raise ValueError("payload exceeds capacity at every threshold")
    else:
        T = _check_T(T, bd)
        marked, lm, info = run(T)
        if info[7] == PEEB_E_CAPACITY:
            raise ValueError(f"payload of {n_bits} bits exceeds capacity {int(info[2])} at T={T}")
    d = _info_dict(info)
    d.pop("status")
    return marked, lm, d


def pee_extract(marked, lm_packed, T, n_bits, bit_depth=None, device=None, predictor="rhombus"):
    """-> (payload packed uint8[ceil(n_bits/8)], recovered image)."""
    marked = _cabi.as_image(marked, "marked")
    if marked.ndim != 2:
        raise ValueError("marked must be 2-D")
    lm = np.ascontiguousarray(lm_packed, dtype=np.uint8)
    payload, rec, info = pee_extract_batch(marked[None], lm[None], T, [int(n_bits)], bit_depth, device=device,
                                           predictor=predictor)
    if info[0, 7] == PEEB_E_CAPACITY:
        raise ValueError("n_bits exceeds the number of carriers found")
    return payload[0, :(int(n_bits) + 7) // 8], rec[0]


def pee_sweep(img, payload, T_values, bit_depth=None, n_bits=None, device=None):
    """Capacity / distortion table, one real embed per T (the (image, T) pair is
    the unit of work; the cover is shared and no marked image is written).
    -> list of dicts {T, capacity, cap0, cap1, n_flagged, sse, mse, psnr}."""
    img = _cabi.as_image(img)
    if img.ndim != 2:
        raise ValueError("image must be 2-D")
    bd = _bit_depth(img, bit_depth)
    maxval = (1 << bd) - 1
    packed, n_bits = pack_payload(payload, n_bits)
    Ts = np.asarray(list(T_values), dtype=np.int32)
    n = Ts.size
    if n == 0:
        return []
    _, _, info = pee_embed_batch(img, packed, [n_bits] * n, Ts, bd, shared_cover=True, shared_payload=True,
                                 want_marked=False, want_lm=False, device=device)
    rows = []
    for k in range(n):
        sse = int(info[k, 6])
        mse = sse / img.size if img.size else 0.0
        psnr = float("inf") if sse == 0 else float(10 * np.log10((maxval ** 2) / mse))
        rows.append({"T": int(Ts[k]), "capacity": int(info[k, 2]), "cap0": int(info[k, 3]), "cap1": int(info[k, 4]),
                     "n_flagged": int(info[k, 5]), "sse": sse, "mse": mse, "psnr": psnr})
    return rows


def pee_sweep_pairs(imgs, payloads, image_index, T_values, bit_depth=None, n_bits=None, device=None):
    """Capacity / distortion of explicit (image, T) pairs -- the sharded unit of the threshold sweep
    over a series (``shard.partition_grid`` hands each rank its pairs).

    imgs (n, h, w); payloads (n, stride) packed; image_index / T_values: equal-length int sequences,
    grouped by image (any order works, consecutive equal indices share one launch).
    -> int64 array (len(pairs), 8): the ``info`` rows {T, n_bits, capacity, cap0, cap1, n_flagged, sse, status}.
    """
    imgs = _cabi.as_image(imgs, "imgs")
    if imgs.ndim != 3:
        raise ValueError("imgs must be (n, h, w)")
    bd = _bit_depth(imgs, bit_depth)
    payloads = np.ascontiguousarray(payloads, dtype=np.uint8)
    idx = np.asarray(image_index, dtype=np.int64).reshape(-1)
    Ts = np.asarray(T_values, dtype=np.int32).reshape(-1)
    if idx.size != Ts.size:
        raise ValueError("image_index and T_values must have the same length")
    out = np.zeros((idx.size, INFO), np.int64)
    nb_all = 8 * payloads.shape[1] if n_bits is None else int(n_bits)
    k = 0
    while k < idx.size:
        j = k
        while j < idx.size and idx[j] == idx[k]:
            j += 1
        u = int(idx[k])
        _, _, info = pee_embed_batch(imgs[u], payloads[u], [nb_all] * (j - k), Ts[k:j], bd, shared_cover=True,
                                     shared_payload=True, want_marked=False, want_lm=False, device=device)
        out[k:j] = info
        k = j
    return out
