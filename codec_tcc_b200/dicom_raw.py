"""N3: pixel data of *uncompressed* DICOM files without pydicom, so that series
can be fed to the GPU path straight from disk into pinned host memory.

Replaces, for the two uncompressed little-endian transfer syntaxes, what the reference does with
``pydicom.dcmread(path).pixel_array`` (src/codec.py:211-213, src/mse.py:18-37): frame 0 of a
multi-frame object, ``int16`` reinterpreted as ``uint16``, ``BitsStored`` giving the value range.
Compressed transfer syntaxes (JPEG-LS, JPEG 2000, deflate, RLE ...) raise ``ValueError`` -- those
need the codecs the reference shells out to (out of scope).

Host-only code: a tag walk over the data set; no pixel arithmetic happens here.
"""
from __future__ import annotations

import struct

import numpy as np

IMPLICIT_LE = "1.2.840.10008.1.2"
EXPLICIT_LE = "1.2.840.10008.1.2.1"
_LONG_VR = {b"OB", b"OD", b"OF", b"OL", b"OV", b"OW", b"SQ", b"UC", b"UN", b"UR", b"UT"}
_WANTED = {
    (0x0028, 0x0002): "SamplesPerPixel", (0x0028, 0x0008): "NumberOfFrames", (0x0028, 0x0010): "Rows",
    (0x0028, 0x0011): "Columns", (0x0028, 0x0100): "BitsAllocated", (0x0028, 0x0101): "BitsStored",
    (0x0028, 0x0103): "PixelRepresentation",
}


def _skip_undefined_sequence(buf, pos, explicit):
    """pos = first byte after an element header with undefined length; returns the offset after the
    matching sequence delimiter (FFFE,E0DD).  Items may nest sequences of undefined length."""
    depth = 1
    while depth and pos + 8 <= len(buf):
        group, elem = struct.unpack_from("<HH", buf, pos)
        if group == 0xFFFE:  # item / item delimiter / sequence delimiter: always 4-byte length, no VR
            length = struct.unpack_from("<I", buf, pos + 4)[0]
            pos += 8
            if elem == 0xE0DD:
                depth -= 1
            elif elem == 0xE000 and length != 0xFFFFFFFF:
                pos += length
            continue
        vr, length, hdr = _element_header(buf, pos, explicit)
        pos += hdr
        if length == 0xFFFFFFFF:
            depth += 1
        else:
            pos += length
    return pos


def _element_header(buf, pos, explicit):
    if explicit:
        vr = bytes(buf[pos + 4:pos + 6])
        if vr in _LONG_VR:
            return vr, struct.unpack_from("<I", buf, pos + 8)[0], 12
        return vr, struct.unpack_from("<H", buf, pos + 6)[0], 8
    return b"", struct.unpack_from("<I", buf, pos + 4)[0], 8


def parse(path):
    """-> dict(info) with the image geometry, transfer syntax and ``pixel_offset`` / ``pixel_length``
    (byte range of (7FE0,0010) in the file)."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 132 or buf[128:132] != b"DICM":
        raise ValueError(f"{path}: not a DICOM part-10 file (no DICM marker)")
    pos, ts = 132, IMPLICIT_LE
    # file meta group (0002,xxxx) is always explicit VR little endian
    while pos + 8 <= len(buf):
        group, elem = struct.unpack_from("<HH", buf, pos)
        if group != 0x0002:
            break
        vr, length, hdr = _element_header(buf, pos, True)
        if elem == 0x0010:
            ts = bytes(buf[pos + hdr:pos + hdr + length]).rstrip(b"\x00 ").decode("ascii")
        pos += hdr + length
    if ts not in (IMPLICIT_LE, EXPLICIT_LE):
        raise ValueError(f"{path}: transfer syntax {ts} is not an uncompressed little-endian one; "
                         "decode it with the reference's codecs (src/codec.py:167-209) first")
    explicit = ts == EXPLICIT_LE
    info = {"TransferSyntaxUID": ts, "SamplesPerPixel": 1, "NumberOfFrames": 1, "PixelRepresentation": 0}
    while pos + 8 <= len(buf):
        group, elem = struct.unpack_from("<HH", buf, pos)
        vr, length, hdr = _element_header(buf, pos, explicit)
        if (group, elem) == (0x7FE0, 0x0010):
            if length == 0xFFFFFFFF:
                raise ValueError(f"{path}: encapsulated (compressed) pixel data")
            info["pixel_offset"], info["pixel_length"] = pos + hdr, length
            break
        pos += hdr
        if length == 0xFFFFFFFF:
            pos = _skip_undefined_sequence(buf, pos, explicit)
            continue
        key = _WANTED.get((group, elem))
        if key:
            raw = bytes(buf[pos:pos + length])
            if key == "NumberOfFrames":  # IS: integer string
                info[key] = int(raw.decode("ascii").strip("\x00 ") or 1)
            else:                          # US
                info[key] = struct.unpack_from("<H", raw)[0]
        pos += length
    for need in ("Rows", "Columns", "BitsAllocated", "pixel_offset"):
        if need not in info:
            raise ValueError(f"{path}: element {need} not found")
    info.setdefault("BitsStored", info["BitsAllocated"])
    return info


def read_pixels(path, out=None):
    """Frame 0 as a C-contiguous ``(rows, cols)`` ``uint8`` / ``uint16`` array plus the info dict
    (``BitsStored`` gives ``maxval = 2**BitsStored - 1`` for the PEE path).  ``out``: optional
    destination (e.g. a row of a pinned batch from ``_cabi.pinned_empty``)."""
    info = parse(path)
    if info["SamplesPerPixel"] != 1:
        raise ValueError(f"{path}: {info['SamplesPerPixel']} samples per pixel; the codec path is grayscale (src/codec.py:34)")
    if info["BitsAllocated"] not in (8, 16):
        raise ValueError(f"{path}: BitsAllocated {info['BitsAllocated']} unsupported")
    dt = np.dtype("<u2") if info["BitsAllocated"] == 16 else np.dtype("u1")
    h, w = info["Rows"], info["Columns"]
    need = h * w * dt.itemsize
    if info["pixel_length"] < need:
        raise ValueError(f"{path}: pixel data shorter than one frame")
    frame = np.memmap(path, dtype=dt, mode="r", offset=info["pixel_offset"], shape=(h, w))
    if out is None:
        out = np.empty((h, w), dt.newbyteorder("="))
    out[...] = frame  # signed 16-bit data keeps its bit pattern, like the reference's astype(np.uint16) (src/mse.py:28-29)
    return out, info


def carregar_imagem(path):
    """``AnalisadorMSE.carregar_imagem`` for uncompressed .dcm files (src/mse.py:18-37):
    ``(float64 array, max_valor, bits_stored)``."""
    arr, info = read_pixels(path)
    bits = info["BitsStored"]
    return arr.astype(np.float64), (1 << bits) - 1, bits
