"""The pixel-array functions of the reference's ``src/codec.py`` (rows a5-a9 of
SURVEY.md section 8) with the same names, arguments, return values and error
behaviour, executed by hand-written sm_100a kernels through the C ABI of
``include/peeb200.h``.

What stays on the host is only what the reference itself does per call in
O(s) or on scalars: the segment plan (Python's ``random.shuffle`` with seed 42
must be CPython's), the float64 entropy sums over the *histogram* (so that
numpy's summation order -- and therefore the split point ``s`` -- is reproduced
bit for bit), the exact-rational arg-max over per-tile moments, and the final
bytes->utf-8 decoding.  There is no CPU implementation of the per-pixel work.

DICOM / JPEG-XL / container I/O (src/codec.py:19-213, 601-750, 795-926) is out
of scope.
"""
from __future__ import annotations

import ctypes as C
import math
import random
from fractions import Fraction

import numpy as np

from . import _cabi
from ._cabi import check, lib, ptr, workspace

__all__ = [
    "message_to_bits", "distribute_message_segments", "calculate_entropy", "calculate_mutual_information",
    "adaptive_modalities_decomposition", "merge_modalities", "extract_local_planes", "lsb_embed_multi_plane",
    "lsb_embed_block_then_multiplane", "decode_message", "embed_pipeline",
    "hybrid_start_offset", "extraction_plan", "recover_cover", "extract_message_bits", "extract_message",
]

VERBOSE = False  # the reference prints from inside its numerics (src/codec.py:568,577-578); opt in to that


# ------------------------------------------------------------------ host-side helpers
def message_to_bits(message: str) -> str:
    """src/codec.py:239-240: '0'/'1' string, 8 bits per character, MSB first."""
    return "".join(format(ord(ch), "08b") for ch in message)


def distribute_message_segments(local_planes, message_bits):
    """src/codec.py:242-274 -> ``(segments, distributed_sizes, segment_indices)``.
    O(s) host logic; reseeds Python's global ``random`` with 42 exactly like the
    reference (src/codec.py:263)."""
    s = len(local_planes)
    total_bits = len(message_bits)
    weights = [(s - i) ** 2 for i in range(s)]
    weight_sum = sum(weights)
    sizes = [max(1, int((wt / weight_sum) * total_bits)) for wt in weights]
    surplus = sum(sizes) - total_bits
    if surplus != 0:
        sizes[sizes.index(max(sizes))] -= surplus
    order = list(range(s))
    random.seed(42)
    random.shuffle(order)
    segments, at = [], 0
    for plane in order:
        segments.append(message_bits[at:at + sizes[plane]])
        at += sizes[plane]
    return segments, sizes, order


def _bits_to_packed(bits: str) -> np.ndarray:
    """'0'/'1' string -> packed uint8 (MSB first)."""
    if not bits:
        return np.zeros(0, np.uint8)
    raw = np.frombuffer(bits.encode("ascii"), dtype=np.uint8)
    vals = raw - ord("0")
    if vals.max() > 1:
        raise ValueError("invalid literal for int() with base 10: message bits must be '0'/'1'")
    return np.packbits(vals)


def _entropy_from_counts(counts, total):
    """numpy's own order of operations (src/codec.py:498-501), on a histogram."""
    p = counts[counts > 0] / total
    return -np.sum(p * np.log2(p))


def _plane_stack(planes, what):
    """list of equally shaped planes -> (list of C-contiguous arrays, shape, dtype)."""
    if len(planes) == 0:
        raise IndexError("list index out of range")  # the reference indexes planes[0]
    arrs = [np.asarray(p) for p in planes]
    dt = arrs[0].dtype
    if dt not in (np.uint8, np.uint16):
        raise ValueError(f"{what} must be uint8 or uint16 planes, got {dt}")
    shape = arrs[0].shape
    out = []
    for a in arrs:
        if a.shape != shape:
            raise ValueError(f"{what}: all planes must share one shape")
        out.append(np.ascontiguousarray(a if a.dtype == dt else a.astype(dt)))
    return out, shape, dt


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


# ------------------------------------------------------------------ a5: entropy / MI / split
def _image_histogram(image, device=None):
    """(hist int64[nbins], plane_ones int64[16]) from the device."""
    img = _cabi.as_image(image, "image_array")
    hist = np.zeros(65536, np.uint32)
    ones = np.zeros(16, np.uint64)
    ws = workspace(device)
    check(lib().peeb_hist_planes_h(ws.handle, ptr(img), img.size, img.dtype.itemsize, ptr(hist), ptr(ones)),
          "peeb_hist_planes_h")
    nbins = 256 if img.dtype == np.uint8 else 65536
    return hist[:nbins].astype(np.int64), ones.astype(np.int64)


def calculate_entropy(data_array):
    """src/codec.py:489-502.  The histogram comes from the GPU; ``np.bincount``
    stops at the largest value present, so trailing empty bins are cut before
    the (order-sensitive) float sum."""
    arr = _cabi.as_image(data_array, "data_array")
    if arr.size == 0:
        return -np.sum(np.zeros(0))  # the reference yields -0.0 for an empty array
    hist, _ = _image_histogram(arr)
    return _entropy_from_counts(hist, arr.size)


def _plane_information(hist, ones, total, bit, nbins):
    """I(plane; image) from the image histogram and the plane's population
    count -- the three entropies of src/codec.py:529-554, in numpy's order.  The
    non-empty joint bins are ``hist``'s non-empty bins, those whose value has
    the bit clear first, then those with the bit set (index = bit*(max+1)+v,
    src/codec.py:548)."""
    zeros = total - ones
    if ones == 0 or zeros == 0 or np.count_nonzero(hist) <= 1:
        return 0.0  # src/codec.py:520-523
    h_x = _entropy_from_counts(np.array([zeros, ones], dtype=np.int64), total)
    h_y = _entropy_from_counts(hist, total)
    has_bit = ((np.arange(nbins) >> bit) & 1).astype(bool)
    joint = np.concatenate([np.where(has_bit, 0, hist), np.where(has_bit, hist, 0)])
    h_xy = _entropy_from_counts(joint, total)
    return max(0.0, h_x + h_y - h_xy)


def calculate_mutual_information(bit_plane, image_array):
    """src/codec.py:504-559: I(X;Y) = H(X) + H(Y) - H(X,Y) of a plane X and the image Y.

    A plane that is one of the image's own bit planes (how the reference calls it, :571,:588) needs no joint
    histogram: the image histogram and the plane's population count give all three entropies.  Which bit it is
    gets identified on the device: the population count narrows the candidates, an exact SSE of zero against
    the unpacked candidate confirms.  Any other plane (the reference accepts whatever ``np.bincount`` does) takes
    one more device histogram per distinct plane value: the image with the pixels outside that value zeroed,
    bin 0 corrected by their number -- the joint counts in ``np.bincount``'s order (index = x * (max + 1) + y,
    src/codec.py:546-548)."""
    img = _cabi.as_image(image_array, "image_array")
    plane = np.ascontiguousarray(np.asarray(bit_plane))
    if plane.shape != img.shape:
        raise ValueError("bit_plane and image_array must have the same shape")
    if plane.dtype not in (np.uint8, np.uint16):
        if not np.issubdtype(plane.dtype, np.integer) and plane.dtype != np.bool_:
            raise TypeError(f"Cannot cast array data from {plane.dtype} to dtype('int64') according to the rule 'safe'")
        if plane.size and (plane.min() < 0 or plane.max() > 65535):
            raise ValueError("bit_plane values must be in 0..65535")
        plane = plane.astype(np.uint16)
    if img.size == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    hist, ones = _image_histogram(img)
    nbins = hist.size
    total = img.size
    from .mse import image_moments
    pm = image_moments(plane, plane)
    if pm["max_a"] <= 1:
        cmp_plane = plane if plane.dtype == img.dtype else plane.astype(img.dtype)
        for bit in range(8 * img.dtype.itemsize):
            if int(ones[bit]) != pm["sum_a"]:
                continue
            cand = extract_bit_plane(img, bit)
            if image_moments(cand, cmp_plane)["sse"] == 0:
                return _plane_information(hist, int(ones[bit]), total, bit, nbins)
    # general plane: counts of its values, then one masked image histogram per value
    xhist, _ = _image_histogram(plane)
    xvals = np.flatnonzero(xhist)
    if xvals.size <= 1 or np.count_nonzero(hist) <= 1:
        return 0.0  # src/codec.py:520-523
    h_x = _entropy_from_counts(xhist, total)
    h_y = _entropy_from_counts(hist, total)
    joint = []
    for v in xvals:
        sel = plane == v
        masked, _ = _image_histogram(np.where(sel, img, img.dtype.type(0)))
        masked[0] -= total - int(xhist[v])
        joint.append(masked)
    h_xy = _entropy_from_counts(np.concatenate(joint), total)
    return max(0.0, h_x + h_y - h_xy)


def extract_bit_plane(image, bit, device=None):
    """(image >> bit) & 1 with the image's dtype (src/codec.py:571)."""
    img = _cabi.as_image(image, "image")
    out = _cabi.out_empty((1,) + img.shape, img.dtype)
    ws = workspace(device)
    check(lib().peeb_planes_unpack_h(ws.handle, ptr(img), img.size, img.dtype.itemsize, int(bit), 1, ptr(out)),
          "peeb_planes_unpack_h")
    return out[0]


def choose_split(image_array, beta=0.8, nbits=None):
    """The decision of src/codec.py:573-593 -> ``(s, total_info, [mi ...])``."""
    img = _cabi.as_image(image_array, "image_array")
    nbits = img.dtype.itemsize * 8 if nbits is None else nbits
    hist, ones = _image_histogram(img)
    total = img.size
    total_info = _entropy_from_counts(hist, total)
    target = beta * total_info
    acc, s, seen = 0.0, 1, []
    for i in range(nbits):
        o = int(ones[i]) if i < 16 else 0
        mi = _plane_information(hist, o, total, i, hist.size) if i < 8 * img.dtype.itemsize else 0.0
        seen.append(mi)
        acc += mi
        if acc >= target:
            s = i + 1
            break
    return s, total_info, seen


def adaptive_modalities_decomposition(image_array, beta=0.8, nbits=None):
    """src/codec.py:561-599 -> ``(global_planes, local_planes)``: lists of (h, w)
    0/1 arrays with the image's dtype, split at the first ``s`` where the
    cumulative plane information reaches ``beta * H(image)``."""
    img = _cabi.as_image(image_array, "image_array")
    nbits = img.dtype.itemsize * 8 if nbits is None else nbits
    if VERBOSE:
        print(f"   - Profundidade de bits efetiva: {nbits}")
    s, total_info, _ = choose_split(img, beta, nbits)
    if VERBOSE:
        print(f"   - Informação total da imagem: {total_info:.4f}")
        print(f"   - Meta de retenção ({beta*100}%): {beta * total_info:.4f}")
    planes = _cabi.out_empty((max(nbits, 0),) + img.shape, img.dtype)
    if nbits > 0 and img.size:
        ws = workspace()
        check(lib().peeb_planes_unpack_h(ws.handle, ptr(img), img.size, img.dtype.itemsize, 0, nbits, ptr(planes)),
              "peeb_planes_unpack_h")
    bit_planes = [planes[i] for i in range(nbits)]
    return bit_planes[s:], bit_planes[:s]


# ------------------------------------------------------------------ a8: pack / unpack
def merge_modalities(global_planes, local_planes):
    """src/codec.py:215-237: OR of ``plane.astype(dtype) << k``; dtype uint16
    when there are more than 8 planes in total, else uint8."""
    allp = list(local_planes) + list(global_planes)
    if not allp:
        raise IndexError("list index out of range")
    planes, shape, dt = _plane_stack(allp, "planes")
    total_bits = len(planes)
    out_dtype = np.uint16 if total_bits > 8 else np.uint8
    out = _cabi.out_empty(shape, out_dtype)
    use = planes[:16]  # a uint16 shifted by >= 16 contributes nothing
    n = int(np.prod(shape)) if len(shape) else 1
    if n == 0:
        return np.zeros(shape, out_dtype)
    ws = workspace()
    if total_bits > 16:
        # keep the uint16 output decision while packing only the first 16 planes
        tmp = _cabi.out_empty(shape, np.uint16)
        check(lib().peeb_planes_pack_h(ws.handle, _ptr_array(use), n, dt.itemsize, 16, ptr(tmp)), "peeb_planes_pack_h")
        return tmp
    check(lib().peeb_planes_pack_h(ws.handle, _ptr_array(use), n, dt.itemsize, len(use), ptr(out)), "peeb_planes_pack_h")
    return out


def extract_local_planes(stego_array, s):
    """src/codec.py:789-793: the ``s`` least significant bit planes."""
    img = _cabi.as_image(stego_array, "stego_array")
    s = int(s)
    planes = _cabi.out_empty((max(s, 0),) + img.shape, img.dtype)
    if s > 0 and img.size:
        ws = workspace()
        check(lib().peeb_planes_unpack_h(ws.handle, ptr(img), img.size, img.dtype.itemsize, 0, s, ptr(planes)),
              "peeb_planes_unpack_h")
    return [planes[i] for i in range(s)]


# ------------------------------------------------------------------ a6: tile variance arg-max
def _dyadic_mean(k, n):
    """True when the float evaluation of np.var on n values summing to k (all
    0/1) is exact, so equal rationals give equal floats."""
    if k == 0 or k == n:
        return True
    r = n // math.gcd(k, n)
    return (r & (r - 1)) == 0


def _best_tile_offset(ref_plane, sbs, device=None):
    """Raster offset ``y*w + x`` of the first tile attaining the largest
    ``float(np.var(tile))`` (src/codec.py:437-453, strict '>' so the first wins).
    The device returns sum and sum of squares per tile (row-major over tiles)."""
    plane = np.ascontiguousarray(ref_plane)
    h, w = plane.shape
    sbs = int(sbs)
    if sbs < 1:
        raise ValueError("range() arg 3 must not be zero" if sbs == 0 else "search_block_size must be positive")
    ty, tx = -(-h // sbs), -(-w // sbs)
    sums = np.zeros((ty * tx, 2), np.int64)
    ws = workspace(device)
    check(lib().peeb_tile_moments_h(ws.handle, ptr(plane), h, w, plane.dtype.itemsize, sbs, ptr(sums)),
          "peeb_tile_moments_h")
    return _tile_argmax_from_moments(plane, sbs, sums)


def _tile_argmax_from_moments(plane, sbs, sums, shape=None):
    """Host half of the tile search: var = (n*sq - s^2)/n^2 per tile is compared
    exactly (rationals).  When exact ties (or near ties) involve a tile whose
    float evaluation is not exact, the reference's float result decides, so
    ``float(np.var(tile))`` is evaluated for those few tiles only.  ``plane`` may
    be a zero-argument callable that fetches the plane only in that case."""
    h, w = shape if shape is not None else plane.shape
    ty, tx = -(-h // sbs), -(-w // sbs)
    th = np.minimum(sbs, h - np.arange(ty) * sbs)
    tw = np.minimum(sbs, w - np.arange(tx) * sbs)
    npx = (th[:, None] * tw[None, :]).reshape(-1).astype(np.int64)
    s1, s2 = sums[:, 0], sums[:, 1]
    if float(npx.max()) * float(s2.max() if s2.size else 0) < 2.0 ** 52:
        approx = (npx.astype(np.float64) * s2 - s1.astype(np.float64) ** 2) / (npx.astype(np.float64) ** 2)
    else:
        # n * sum(v^2) no longer fits a float64 mantissa (16-bit planes with large tiles): the difference would cancel
        # to less than the shortlist's margin, so the scores are the correctly rounded exact rationals
        approx = np.array([float(Fraction(int(n) * int(q) - int(a) ** 2, int(n) ** 2)) for n, a, q in zip(npx, s1, s2)])
    top = approx.max()
    short = np.flatnonzero(approx >= top - abs(top) * 1e-9 - 1e-300)
    # exact values once per distinct (n, sum, sum of squares)
    trip, inv = np.unique(np.stack([npx[short], s1[short], s2[short]], axis=1), axis=0, return_inverse=True)
    vals = [Fraction(int(n) * int(q) - int(a) ** 2, int(n) ** 2) for n, a, q in trip]
    best = max(vals)
    is_best = np.array([v == best for v in vals])[inv.reshape(-1)]
    binary = bool(np.all(s1[short] == s2[short]))  # 0/1 tiles: sum == sum of squares
    exact_float = all(_dyadic_mean(int(a), int(n)) for (n, a, q), v in zip(trip, vals) if v == best)
    if binary and bool(is_best.all()) and exact_float:
        pick = int(short[0])
    else:
        arr = plane() if callable(plane) else plane
        pick, best_f = None, -1.0
        for t in (int(v) for v in short):  # raster order over the short list
            y, x = (t // tx) * sbs, (t % tx) * sbs
            f = float(np.var(arr[y:y + sbs, x:x + sbs]))
            if f > best_f:
                best_f, pick = f, t
    return (pick // tx) * sbs * w + (pick % tx) * sbs


# ------------------------------------------------------------------ a6 / a7: embedders
def _embed(local_planes, message_bits, start_offset, advance, device=None):
    planes, shape, dt = _plane_stack(local_planes, "local_planes")
    s = len(planes)
    if len(shape) != 2:
        raise ValueError("planes must be 2-D")
    h, w = shape
    npx = h * w
    segments, sizes, order = distribute_message_segments(planes, message_bits)
    start = np.zeros(s, np.int64)
    length = np.zeros(s, np.int64)
    bit_off = np.zeros(s, np.int64)
    chunks, at, total_used = [], 0, 0
    for seg, plane_idx in zip(segments, order):
        nb = min(len(seg), npx)
        packed = _bits_to_packed(seg[:nb])
        start[plane_idx] = start_offset if npx else 0
        length[plane_idx] = nb
        bit_off[plane_idx] = 8 * at
        chunks.append(packed)
        at += packed.size
        total_used += nb
        if advance and npx:
            start_offset = (start_offset + nb) % npx
    payload = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    out_planes = _cabi.out_empty((s, h, w), dt)
    bitmaps = _cabi.out_empty((s, h, w), np.uint8)
    if npx:
        ws = workspace(device)
        check(lib().peeb_lsb_embed_h(ws.handle, _ptr_array(planes), npx, dt.itemsize, s, ptr(start), ptr(length),
                                     ptr(bit_off), ptr(payload) if payload.size else None, 8 * payload.size,
                                     ptr(out_planes), ptr(bitmaps)), "peeb_lsb_embed_h")
    stego = [out_planes[i] for i in range(s)]
    maps = [bitmaps[i] for i in range(s)]
    return stego, maps, total_used, [int(v) for v in length], sizes, order


def lsb_embed_multi_plane(local_planes, message_bits):
    """src/codec.py:276-318 -> ``(stego_planes, bitmaps, total_used,
    segments_lengths, segment_indices)``; every plane is written from raster
    position 0 and ``segments_lengths`` are the embedded lengths (:315)."""
    stego, maps, used, lens, _sizes, order = _embed(local_planes, message_bits, 0, False)
    return stego, maps, used, lens, order


def lsb_embed_block_then_multiplane(local_planes, message_bits, search_block_size=8,
                                    align_across_planes: bool = False):
    """src/codec.py:412-487: start at the raster offset of the highest-variance
    ``search_block_size`` tile of plane 0, wrap around the image end, advance
    the start by each embedded length unless ``align_across_planes``.
    ``segments_lengths`` are the *planned* sizes (:425,:487)."""
    if len(local_planes) == 0:
        raise IndexError("list index out of range")
    ref = np.asarray(local_planes[0])
    if ref.dtype not in (np.uint8, np.uint16):
        raise ValueError(f"local_planes must be uint8 or uint16 planes, got {ref.dtype}")
    start = _best_tile_offset(ref, search_block_size) if ref.size else 0
    stego, maps, used, _lens, sizes, order = _embed(local_planes, message_bits, start, not align_across_planes)
    return stego, maps, used, sizes, order


# ------------------------------------------------------------------ a9: decode_message
def _decode_bit_chunks(stego_planes, bitmaps, metadata, device=None):
    """Per plane index: (packed bits, count) of the LSBs at the first
    ``segments_lengths[plane]`` positions where the bitmap is non-zero
    (src/codec.py:761-772)."""
    s = metadata["s"]
    chunks = [(np.zeros(0, np.uint8), 0)] * s
    ws = workspace(device)
    for plane_idx in metadata["segments_indices"]:
        plane = np.ascontiguousarray(np.asarray(stego_planes[plane_idx]).reshape(-1))
        if plane.dtype not in (np.uint8, np.uint16):
            raise ValueError("stego planes must be uint8/uint16")
        bm = np.ascontiguousarray(np.asarray(bitmaps[plane_idx]).reshape(-1))
        if bm.dtype != np.uint8:
            bm = (bm != 0).astype(np.uint8)
        if bm.size != plane.size:
            raise IndexError("bitmap and plane sizes differ")
        limit = max(0, min(int(metadata["segments_lengths"][plane_idx]), plane.size))
        bits = np.zeros(((limit + 7) // 8 + 3) // 4 * 4 + 4, np.uint8)
        count = np.zeros(1, np.int64)
        if plane.size and limit:
            check(lib().peeb_compact_bits_h(ws.handle, ptr(plane), ptr(bm), plane.size, plane.dtype.itemsize, limit,
                                            ptr(bits), ptr(count)), "peeb_compact_bits_h")
        chunks[plane_idx] = (bits, int(count[0]))
    return chunks


def decode_message(stego_planes, bitmaps, metadata):
    """src/codec.py:752-787, bug-compatible: reads LSBs only
    where the bitmap is non-zero and concatenates the per-plane pieces by plane
    index; whole bytes are decoded as utf-8 with replacement."""
    chunks = _decode_bit_chunks(stego_planes, bitmaps, metadata)
    pieces = [np.unpackbits(bits)[:cnt] for bits, cnt in chunks]
    allbits = np.concatenate(pieces) if pieces else np.zeros(0, np.uint8)
    nbytes = allbits.size // 8
    raw = np.packbits(allbits[:nbytes * 8]).tobytes()
    return raw.decode("utf-8", errors="replace")


# ------------------------------------------------------------------ N4: the true inverse (opt-in, not in the reference)
def hybrid_start_offset(plane0, search_block_size=8):
    """The raster offset ``lsb_embed_block_then_multiplane`` starts from (src/codec.py:441-453).
    The reference computes it but never returns it; an exact extraction needs it."""
    ref = np.asarray(plane0)
    return _best_tile_offset(ref, search_block_size) if ref.size else 0


def extraction_plan(metadata, npx):
    """Per plane (start, length, stream bit offset) and the total, as the embedders laid the message
    out (src/codec.py:288-316 / :456-485).  ``metadata``: ``s``, ``segments_lengths`` (per plane),
    ``segments_indices`` (embed order), and for the hybrid embedder ``hybrid=True``, ``start_offset``,
    ``align_across_planes``."""
    s = int(metadata["s"])
    start_offset = int(metadata.get("start_offset", 0))
    advance = bool(metadata.get("hybrid", False)) and not bool(metadata.get("align_across_planes", False))
    start, length, off = np.zeros(s, np.int64), np.zeros(s, np.int64), np.zeros(s, np.int64)
    # The hybrid embedder reports the *planned* sizes (src/codec.py:425,487); a message shorter than the
    # plan (sizes are at least 1 per plane, :253) leaves the last segments short or empty.  With
    # ``message_bits`` (the length of the embedded bit string) the actual lengths follow from the slicing
    # at :267-272.
    mbits = metadata.get("message_bits")
    at, mpos = 0, 0
    for plane_idx in metadata["segments_indices"]:
        size = int(metadata["segments_lengths"][plane_idx])  # may even be negative for a message shorter than s bits
        if mbits is None:
            seglen = max(0, size)
        else:  # the very slice of src/codec.py:269-271, Python semantics included
            lo, hi, _ = slice(mpos, mpos + size).indices(int(mbits))
            seglen = max(0, hi - lo)
        mpos += size
        nb = min(seglen, npx)
        start[plane_idx], length[plane_idx], off[plane_idx] = (start_offset if npx else 0), nb, at
        at += nb
        if advance and npx:
            start_offset = (start_offset + nb) % npx
    return start, length, off, at


def recover_cover(stego_array, bitmaps, device=None):
    """The cover image from a stego image and the XOR side bitmaps of ``lsb_embed_*``
    (src/codec.py:309-311 inverted; the reference's ``decode_bin`` never does this):
    ``cover = stego ^ sum_p (bitmap_p << p)``."""
    img = _cabi.as_image(stego_array, "stego_array")
    maps = [np.ascontiguousarray(np.asarray(b).reshape(-1)) for b in bitmaps]
    s = len(maps)
    if s < 1 or s > 8 * img.dtype.itemsize:
        raise ValueError("between 1 and bits-per-pixel bitmaps are needed")
    for b in maps:
        if b.dtype != np.uint8 or b.size != img.size:
            raise ValueError("bitmaps must be uint8 arrays of the image's size")
    out = _cabi.out_empty(img.shape, img.dtype)
    if img.size:
        ws = workspace(device)
        check(lib().peeb_lsb_recover_h(ws.handle, ptr(img), _ptr_array(maps), img.size, img.dtype.itemsize, s, ptr(out)),
              "peeb_lsb_recover_h")
    return out


def extract_message_bits(stego_array, metadata, device=None):
    """The embedded bit string, exactly (the inverse of ``lsb_embed_multi_plane`` /
    ``lsb_embed_block_then_multiplane``; ``decode_message`` above stays bug-compatible)."""
    img = _cabi.as_image(stego_array, "stego_array")
    start, length, off, total = extraction_plan(metadata, img.size)
    out = np.zeros((total + 7) // 8 + 8, np.uint8)
    if total and img.size:
        ws = workspace(device)
        check(lib().peeb_lsb_extract_h(ws.handle, ptr(img), img.size, img.dtype.itemsize, int(metadata["s"]), ptr(start),
                                       ptr(length), ptr(off), total, ptr(out)), "peeb_lsb_extract_h")
    return (np.unpackbits(out)[:total] + np.uint8(48)).tobytes().decode("ascii")  # 0 / 1 -> "0" / "1"


def extract_message(stego_array, metadata, device=None):
    bits = extract_message_bits(stego_array, metadata, device)
    nbytes = len(bits) // 8
    raw = np.packbits(np.frombuffer(bits[:nbytes * 8].encode(), np.uint8) - 48).tobytes() if nbytes else b""
    return raw.decode("utf-8", errors="replace")


# ------------------------------------------------------------------ the reference's encode flow, device resident
def embed_pipeline(image_array, message_bits, beta=0.8, search_block_size=16, align_across_planes=False,
                   hybrid=True, nbits=None, device=None, bitmaps_as=None):
    """Steps 3-5 of the reference's ``main()`` (src/codec.py:868-880) in one call:
    ``adaptive_modalities_decomposition`` -> ``lsb_embed_block_then_multiplane`` (or
    ``lsb_embed_multi_plane`` with ``hybrid=False``) -> ``merge_modalities``, with the bit planes
    living only in device memory: one upload of the image, one download of the stego image and the
    uint8 bitmaps.  Results are identical to chaining the three functions.

    -> ``(stego_image, bitmaps (s, h, w) uint8, meta)`` with ``meta = {'s', 'segments_lengths',
    'segments_indices', 'total_used', 'start_offset'}`` (what ``create_header`` needs, :895-905).

    ``bitmaps_as="pbr"``: the bitmaps come back as the "PBR1" blob ``container.pack_bitmaps(...,
    coding="pbr")`` would make of them (step 7 of ``main()``, src/codec.py:887-889, folded in): they are coded
    where they are, in device memory, and only the blob -- a few kilobytes for a text message instead of
    ``s * h * w`` bytes -- crosses PCIe.  ``container.unpack_bitmaps(blob, s)`` restores the arrays.
    """
    if bitmaps_as not in (None, "pbr"):
        raise ValueError("bitmaps_as must be None or 'pbr'")
    img = _cabi.as_image(image_array, "image_array")
    if img.ndim != 2:
        raise ValueError("A imagem deve ser 2D (grayscale).")  # src/codec.py:34
    nb = img.dtype.itemsize * 8 if nbits is None else int(nbits)
    if nb < 1 or nb > 16:
        raise ValueError("nbits must be 1..16")
    h, w = img.shape
    npx, item = img.size, img.dtype.itemsize
    L, ws = lib(), workspace(device)
    st = ws.stream
    d_img = _cabi.DeviceBuffer(ws, npx * item)
    d_hist = _cabi.DeviceBuffer(ws, 65536 * 4 + 16 * 8)
    d_planes = _cabi.DeviceBuffer(ws, nb * npx * item)
    d_img.upload(img)
    # (1) split point from the device histogram (float64 sums on the host, numpy's order)
    check(L.peeb_hist_planes(ws.handle, d_img.ptr, npx, item, d_hist.ptr, d_hist.ptr + 65536 * 4, st), "peeb_hist_planes")
    hist = d_hist.download(np.zeros(65536, np.uint32))
    ones = d_hist.download(np.zeros(16, np.uint64), 65536 * 4).astype(np.int64)
    nbins = 256 if item == 1 else 65536
    hist = hist[:nbins].astype(np.int64)
    total_info = _entropy_from_counts(hist, npx)
    acc, s = 0.0, 1
    for i in range(nb):
        acc += _plane_information(hist, int(ones[i]), npx, i, nbins) if i < 8 * item else 0.0
        if acc >= beta * total_info:
            s = i + 1
            break
    # (2) all planes on the device; the first s are the local ones
    check(L.peeb_planes_unpack(ws.handle, d_img.ptr, npx, item, 0, nb, d_planes.ptr, st), "peeb_planes_unpack")
    # (3) start offset from the tile moments of local plane 0
    start = 0
    if hybrid:
        sbs = int(search_block_size)
        if sbs < 1:
            raise ValueError("search_block_size must be positive")
        ntiles = (-(-h // sbs)) * (-(-w // sbs))
        d_sums = _cabi.DeviceBuffer(ws, ntiles * 16)
        check(L.peeb_tile_moments(ws.handle, d_planes.ptr, h, w, item, sbs, d_sums.ptr, st), "peeb_tile_moments")
        sums = d_sums.download(np.zeros((ntiles, 2), np.int64))
        start = _tile_argmax_from_moments(lambda: d_planes.download(np.empty((h, w), img.dtype)), sbs, sums, (h, w))
        d_sums.free()
    # (4) segment plan + embed into the local planes (in a second buffer), bitmaps beside
    segments, sizes, order = distribute_message_segments([None] * s, message_bits)
    starts, lens, offs = np.zeros(s, np.int64), np.zeros(s, np.int64), np.zeros(s, np.int64)
    chunks, at, used, cur = [], 0, 0, start
    for seg, plane_idx in zip(segments, order):
        n_seg = min(len(seg), npx)
        packed = _bits_to_packed(seg[:n_seg])
        starts[plane_idx], lens[plane_idx], offs[plane_idx] = cur, n_seg, 8 * at
        chunks.append(packed)
        at += packed.size
        used += n_seg
        if hybrid and not align_across_planes and npx:
            cur = (cur + n_seg) % npx
    payload = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    d_pay = _cabi.DeviceBuffer(ws, payload.size + 16)
    if payload.size:
        d_pay.upload(payload)
    d_out = _cabi.DeviceBuffer(ws, nb * npx * item)   # stego planes (local) followed by the global planes
    d_bm = _cabi.DeviceBuffer(ws, s * npx)
    check(L.peeb_lsb_embed(ws.handle, d_planes.ptr, npx, item, s, ptr(starts), ptr(lens), ptr(offs), d_pay.ptr,
                           8 * payload.size, d_out.ptr, d_bm.ptr, st), "peeb_lsb_embed")
    # (5) merge: local stego planes + untouched global planes -> stego image
    if nb > s:
        check(L.peeb_planes_unpack(ws.handle, d_img.ptr, npx, item, s, nb - s, d_out.ptr + s * npx * item, st),
              "peeb_planes_unpack")
    out_dtype = np.uint16 if nb > 8 else np.uint8
    d_stego = _cabi.DeviceBuffer(ws, npx * out_dtype().itemsize)
    check(L.peeb_planes_pack(ws.handle, d_out.ptr, npx, item, nb, d_stego.ptr, st), "peeb_planes_pack")
    stego = d_stego.download(_cabi.out_empty((h, w), out_dtype))
    if bitmaps_as == "pbr":
        import ctypes as C
        cap = int(L.peeb_bitmap_blob_bound(s * npx))
        d_blob = _cabi.DeviceBuffer(ws, cap)
        got = C.c_int64(0)
        check(L.peeb_bitmap_encode(ws.handle, d_bm.ptr, s * npx, 0, d_blob.ptr, cap, C.byref(got), st), "peeb_bitmap_encode")
        bitmaps = d_blob.download(np.empty(got.value, np.uint8)).tobytes()
        d_blob.free()
    else:
        bitmaps = d_bm.download(_cabi.out_empty((s, h, w), np.uint8))
    for b in (d_img, d_hist, d_planes, d_pay, d_out, d_bm, d_stego):
        b.free()
    meta = {"s": s, "segments_lengths": sizes if hybrid else [int(v) for v in lens],
            "segments_indices": order, "total_used": used, "start_offset": int(start)}
    return stego, bitmaps, meta
