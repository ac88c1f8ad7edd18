"""N2: the reference's ``.bin`` container, byte for byte.

``create_header`` / ``create_binary_file`` / ``parse_bin_file`` (src/codec.py:601-670, 689-750)
with the same arguments, return values and error behaviour -- including the 16-bit fields that
make ``struct.pack`` raise for segments longer than 65 535 bits or sides over 65 535 pixels
(SURVEY.md F3.4).  Plus the two blob steps ``main()`` / ``decode_bin`` do around them:
``zlib.compress(np.stack(bitmaps).tobytes())`` (:888-889) and ``np.split(np.frombuffer(
zlib.decompress(...)), s)`` (:820-821).  The container itself is host code; the stego image codecs
(cjxl / gdcmconv, src/codec.py:108-209) stay out of scope: the compressed image is passed in as bytes.

The blob steps also come in a device form, ``coding="pbr"``: the bitmaps are bit packed and their zero
runs removed by CUDA kernels (format "PBR1", ``csrc/peeb_bitcode.cu``), instead of one byte per pixel
going through zlib on the host.  ``unpack_bitmaps`` recognises either blob; the reference's own
``decode_bin`` only reads the zlib form, which stays the default.
"""
from __future__ import annotations

import os
import struct
import zlib

import ctypes as _C

import numpy as np

from ._cabi import check, lib, ptr, workspace

CODEC_IDS = {"png": 1, "j2k": 2, "jls": 3, "jxl": 4}          # src/codec.py:616
CODEC_NAMES = {v: k for k, v in CODEC_IDS.items()}           # src/codec.py:693
VERBOSE = False  # the reference prints the header fields (src/codec.py:647-655); opt in to that


def create_header(codec, s, segments_lengths, segments_indices, bitmaps_blob_size, width, height, start_offset,
                  align_across_planes):
    """src/codec.py:601-656: ``>BBBBHHH`` (version 1, codec id, s, align flag, width, height,
    start offset) + ``s`` x ``H`` lengths + ``s`` x ``B`` indices + ``I`` blob size."""
    codec_id = CODEC_IDS.get(codec.lower(), 0)
    parts = [1, codec_id, s, 1 if align_across_planes else 0, width, height, start_offset]
    parts.extend(segments_lengths)
    parts.extend(segments_indices)
    parts.append(bitmaps_blob_size)
    packed = struct.pack(f">BBBBHHH{s}H{s}BI", *parts)
    if VERBOSE:
        print(f" HEADER: v1 codec {codec_id} ({codec}) s={s} align={parts[3]} {width}x{height} start={start_offset} "
              f"lengths={list(segments_lengths)} indices={list(segments_indices)}")
    return packed


def create_binary_file(filename, header_bytes, stego_compressed, bitmaps_bytes):
    """src/codec.py:658-670: ``STGC`` + ``>I`` header length + header + bitmaps blob + image; returns the file size."""
    with open(filename, "wb") as f:
        f.write(b"STGC")
        f.write(struct.pack(">I", len(header_bytes)))
        f.write(header_bytes)
        f.write(bitmaps_bytes)
        f.write(stego_compressed)
    return os.path.getsize(filename)


def parse_bin_file(filepath):
    """src/codec.py:689-750 -> ``(metadata, bitmaps_data, stego_image_data)``."""
    with open(filepath, "rb") as f:
        if f.read(4) != b"STGC":
            raise ValueError("Arquivo inválido ou com assinatura incorreta.")  # src/codec.py:698
        header_length = struct.unpack(">I", f.read(4))[0]
        header = f.read(header_length)
        base = ">BBBBHHH"
        at = struct.calcsize(base)
        version, codec_id, s, align_flag, width, height, start_offset = struct.unpack(base, header[:at])
        lengths = list(struct.unpack(f">{s}H", header[at:at + 2 * s]))
        at += 2 * s
        indices = list(struct.unpack(f">{s}B", header[at:at + s]))
        at += s
        blob_size = struct.unpack(">I", header[at:at + 4])[0]
        bitmaps_data = f.read(blob_size)
        stego_image_data = f.read()
    metadata = {"version": version, "codec": CODEC_NAMES.get(codec_id, "unknown"), "s": s, "align_flag": align_flag,
                "width": width, "height": height, "start_offset": start_offset, "segments_lengths": lengths,
                "segments_indices": indices}
    return metadata, bitmaps_data, stego_image_data


PBR_MAGIC = b"PBR1"


def encode_bitmap(elements, packed=False, n=None, device=None) -> bytes:
    """Side bitmap -> "PBR1" blob on the GPU.  ``elements``: uint8 array (non-zero = 1, flattened in C
    order), or with ``packed`` np.packbits bytes holding ``n`` elements (a PEE location map)."""
    a = np.ascontiguousarray(elements)
    if a.dtype != np.uint8:
        raise ValueError("bitmaps must be uint8 arrays")
    a = a.reshape(-1)
    if packed:
        n = a.size * 8 if n is None else int(n)
        if n < 0 or (n + 7) // 8 != a.size:
            raise ValueError("n does not match the packed array's length")
    else:
        n = a.size
    cap = int(lib().peeb_bitmap_blob_bound(n))
    blob = np.empty(cap, np.uint8)
    got = _C.c_int64(0)
    check(lib().peeb_bitmap_encode_h(workspace(device).handle, ptr(a) if a.size else None, n, 1 if packed else 0,
                                     ptr(blob), cap, _C.byref(got)), "peeb_bitmap_encode_h")
    return blob[:got.value].tobytes()


def decode_bitmap(blob, n, packed=False, device=None) -> np.ndarray:
    """"PBR1" blob -> ``n`` uint8 values 0/1 (or the ceil(n/8) np.packbits bytes with ``packed``)."""
    b = np.frombuffer(bytes(blob), np.uint8)
    n = int(n)
    out = np.empty((n + 7) // 8 if packed else n, np.uint8)
    check(lib().peeb_bitmap_decode_h(workspace(device).handle, ptr(b) if b.size else None, b.size,
                                     ptr(out) if out.size else None, n, 1 if packed else 0), "peeb_bitmap_decode_h")
    return out


def pack_bitmaps(bitmaps, coding="zlib", device=None):
    """``zlib.compress(np.stack(bitmaps).tobytes())`` (src/codec.py:888-889); ``coding="pbr"``: the same
    stack of bitmaps bit packed with its zero runs removed, on the GPU."""
    stack = np.stack([np.asarray(b) for b in bitmaps])
    if coding == "zlib":
        return zlib.compress(stack.tobytes())
    if coding != "pbr":
        raise ValueError("coding must be 'zlib' or 'pbr'")
    return encode_bitmap(stack.astype(np.uint8, copy=False), device=device)


def unpack_bitmaps(bitmaps_data, s, device=None):
    """``np.split(np.frombuffer(zlib.decompress(blob), np.uint8), s)`` (src/codec.py:820-821): flat uint8
    arrays; a "PBR1" blob is expanded on the GPU to the same arrays."""
    if bytes(bitmaps_data[:4]) == PBR_MAGIC:
        n = int.from_bytes(bytes(bitmaps_data[8:16]), "little")
        return np.split(decode_bitmap(bitmaps_data, n, device=device), s)
    return np.split(np.frombuffer(zlib.decompress(bitmaps_data), dtype=np.uint8), s)
