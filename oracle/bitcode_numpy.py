"""CPU restatement (TEST INFRASTRUCTURE ONLY) of the "PBR1" side-bitmap coding of
codec_tcc_b200/csrc/peeb_bitcode.cu -- N2 of SURVEY.md 8f.

What it stands in for in the reference: the blob steps around the ``.bin`` container,
``zlib.compress(np.stack(bitmaps).tobytes())`` (src/codec.py:888-889) and
``np.frombuffer(zlib.decompress(blob), np.uint8)`` (src/codec.py:820-821).  The reference has no
bit-packed / run-length format of its own, so the FORMAT is this repository's (parity of the format
is unpinned); what is pinned is the round trip: decode(encode(m)) equals the reference's own
bitmaps on the golden cases (tests/test_bitcode.py).

Layout (little endian):  b"PBR1" | u32 0 | u64 n | u32 nz1 | u32 nz0 | L2 | C0 | non-zero L1 | non-zero L0
  L0[j]  32 elements per word, bytes as np.packbits makes them
  L1[i]  bit t (LSB first) = L0[32 i + t] != 0;   L2[k] likewise over L1
  C0[k]  number of non-zero L0 words among the 1024 under L2[k] (one scan then places every block's words)
Only product code path allowed to import this module: none (tests, smoke and bench checks only).
"""
from __future__ import annotations

import struct

import numpy as np

MAGIC = b"PBR1"
HEADER = 24


def _presence(words: np.ndarray) -> np.ndarray:
    """one bit per 32-bit word (LSB first), packed into 32-bit words"""
    n = words.size
    flags = np.zeros(((n + 31) // 32) * 32, np.uint8)
    flags[:n] = words != 0
    return np.packbits(flags.reshape(-1, 32), axis=1, bitorder="little").view("<u4").reshape(-1)


def encode(elements, packed: bool = False, n: int | None = None) -> bytes:
    """elements: uint8 array, non-zero = 1 (any shape; flattened in C order) -- or, with ``packed``,
    np.packbits bytes holding ``n`` elements."""
    a = np.ascontiguousarray(elements).reshape(-1).view(np.uint8)
    if packed:
        if n is None:
            n = a.size * 8
        bits = np.unpackbits(a)[:n]
    else:
        n = a.size
        bits = (a != 0).astype(np.uint8)
    n0 = (n + 31) // 32
    padded = np.zeros(n0 * 32, np.uint8)
    padded[:n] = bits
    l0 = np.packbits(padded).view("<u4") if n0 else np.zeros(0, "<u4")
    l1 = _presence(l0)
    l2 = _presence(l1)
    nz1, nz0 = l1[l1 != 0], l0[l0 != 0]
    per_block = np.zeros(l2.size * 1024, np.uint32)
    per_block[:l0.size] = l0 != 0
    c0 = per_block.reshape(-1, 1024).sum(axis=1).astype("<u4") if l2.size else np.zeros(0, "<u4")
    return (MAGIC + struct.pack("<IQII", 0, n, nz1.size, nz0.size) + l2.astype("<u4").tobytes() + c0.tobytes()
            + nz1.astype("<u4").tobytes() + nz0.astype("<u4").tobytes())


def _expand(presence: np.ndarray, nonzero: np.ndarray, n_out: int) -> np.ndarray:
    flags = np.unpackbits(presence.astype("<u4").view(np.uint8), bitorder="little")[:n_out].astype(bool)
    if int(flags.sum()) != nonzero.size:
        raise ValueError("corrupt PBR1 blob: level tables disagree")
    out = np.zeros(n_out, "<u4")
    out[flags] = nonzero
    return out


def decode(blob: bytes, n: int, packed: bool = False) -> np.ndarray:
    """-> n uint8 values 0/1, or with ``packed`` the ceil(n/8) np.packbits bytes"""
    if len(blob) < HEADER or blob[:4] != MAGIC:
        raise ValueError("not a PBR1 blob")
    flags, nn, c1, c0 = struct.unpack("<IQII", blob[4:HEADER])
    if flags != 0 or nn != n:
        raise ValueError("PBR1 header mismatch")
    n0 = (n + 31) // 32
    n1 = (n0 + 31) // 32
    n2 = (n1 + 31) // 32
    if len(blob) != HEADER + 4 * (2 * n2 + c1 + c0):
        raise ValueError("PBR1 blob size mismatch")
    body = np.frombuffer(blob, "<u4", offset=HEADER)
    l2, cnt, nz1, nz0 = body[:n2], body[n2:2 * n2], body[2 * n2:2 * n2 + c1], body[2 * n2 + c1:]
    l1 = _expand(l2, nz1, n1)
    l0 = _expand(l1, nz0, n0)
    per_block = np.zeros(n2 * 1024, np.uint32)
    per_block[:n0] = l0 != 0
    if n2 and not np.array_equal(per_block.reshape(-1, 1024).sum(axis=1), cnt):
        raise ValueError("corrupt PBR1 blob: block counts disagree")
    by = l0.astype("<u4").view(np.uint8)
    if packed:
        return by[:(n + 7) // 8].copy()
    return np.unpackbits(by)[:n].copy()
