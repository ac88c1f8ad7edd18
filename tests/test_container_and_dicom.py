"""N2 / N3, host side: the .bin container against the reference's golden header
(Appendix B) and, when the reference is mounted, against its own functions byte for byte; the raw
DICOM reader against the reference's two fixtures (pixel sha256 from Appendix B)."""
import hashlib
import os
import struct

import numpy as np
import pytest

from codec_tcc_b200 import container, dicom_raw

REF = "/root/reference"
HAVE_REF = os.path.exists(os.path.join(REF, "src", "codec.py"))


def test_header_golden_vector():
    # produced by the reference's create_header
    hdr = container.create_header("jxl", 5, [1966, 1256, 706, 314, 78], [3, 1, 2, 4, 0], 1234, 64, 64, 0, False)
    assert len(hdr) == 29 and hdr.hex() == "0104050000400040000007ae04e802c2013a004e0301020400000004d2"


def test_container_round_trip(tmp_path):
    bitmaps = [np.random.default_rng(k).integers(0, 2, (16, 24), dtype=np.uint8) for k in range(3)]
    blob = container.pack_bitmaps(bitmaps)
    hdr = container.create_header("png", 3, [100, 50, 7], [1, 0, 2], len(blob), 24, 16, 77, True)
    path = str(tmp_path / "x.bin")
    size = container.create_binary_file(path, hdr, b"IMAGEBYTES", blob)
    assert size == 4 + 4 + len(hdr) + len(blob) + 10
    meta, bdata, idata = container.parse_bin_file(path)
    assert meta == {"version": 1, "codec": "png", "s": 3, "align_flag": 1, "width": 24, "height": 16, "start_offset": 77,
                    "segments_lengths": [100, 50, 7], "segments_indices": [1, 0, 2]}
    assert idata == b"IMAGEBYTES"
    back = container.unpack_bitmaps(bdata, 3)
    assert all(np.array_equal(a.reshape(-1), b) for a, b in zip(bitmaps, back))


def test_header_limits_match_reference_behaviour():
    with pytest.raises(struct.error):  # 16-bit length fields
        container.create_header("jxl", 1, [70000], [0], 0, 64, 64, 0, False)
    with pytest.raises(ValueError):
        p = "/tmp/_not_stgc.bin"
        open(p, "wb").write(b"NOPE" + b"\0" * 16)
        container.parse_bin_file(p)


@pytest.mark.skipif(not HAVE_REF, reason="reference not mounted")
def test_container_equals_reference(tmp_path):
    from oracle import ref_import
    ref = ref_import.codec()
    args = ("j2k", 4, [163, 91, 40, 10], [2, 1, 3, 0], 4321, 512, 512, 0, False)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        theirs = ref.create_header(*args)
    assert container.create_header(*args) == theirs
    a, b = str(tmp_path / "a.bin"), str(tmp_path / "b.bin")
    blob = container.pack_bitmaps([np.ones((4, 4), np.uint8)] * 4)
    ref.create_binary_file(a, theirs, b"stego", blob)
    container.create_binary_file(b, theirs, b"stego", blob)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert ref.parse_bin_file(a) == container.parse_bin_file(b)


@pytest.mark.skipif(not HAVE_REF, reason="reference images not mounted")
@pytest.mark.parametrize("name,offset,dtype,sha16,bits", [("pe.dcm", 7010, "<u2", "c0903a29144fd600", 12),
                                                         ("torax.dcm", 888, "u1", "a06491169393c35e", 8)])
def test_raw_dicom_reader_on_reference_fixtures(name, offset, dtype, sha16, bits):
    path = os.path.join(REF, "images", name)
    arr, info = dicom_raw.read_pixels(path)
    assert info["pixel_offset"] == offset and arr.shape == (512, 512) and arr.dtype == np.dtype(dtype).newbyteorder("=")
    assert info["BitsStored"] == bits
    assert hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()[:16] == sha16
    f64, maxv, b = dicom_raw.carregar_imagem(path)
    assert f64.dtype == np.float64 and maxv == (1 << bits) - 1 and b == bits and np.array_equal(f64, arr)


def test_raw_dicom_reader_synthetic(tmp_path):
    """A hand-built explicit-VR file with a nested undefined-length sequence before the pixel data."""
    px = np.arange(6 * 8, dtype="<u2").reshape(6, 8) * 37

    def el(g, e, vr, val):
        if vr in (b"OB", b"OW", b"SQ", b"UN"):
            return struct.pack("<HH2sHI", g, e, vr, 0, len(val)) + val
        return struct.pack("<HH2sH", g, e, vr, len(val)) + val

    meta = el(2, 0x10, b"UI", b"1.2.840.10008.1.2.1\0")
    seq = struct.pack("<HH2sHI", 8, 0x1140, b"SQ", 0, 0xFFFFFFFF) + struct.pack("<HHI", 0xFFFE, 0xE000, 0xFFFFFFFF) \
        + el(8, 0x1150, b"UI", b"1.2\0\0") + struct.pack("<HHI", 0xFFFE, 0xE00D, 0) + struct.pack("<HHI", 0xFFFE, 0xE0DD, 0)
    body = seq + el(0x28, 2, b"US", struct.pack("<H", 1)) + el(0x28, 0x10, b"US", struct.pack("<H", 6)) \
        + el(0x28, 0x11, b"US", struct.pack("<H", 8)) + el(0x28, 0x100, b"US", struct.pack("<H", 16)) \
        + el(0x28, 0x101, b"US", struct.pack("<H", 12)) + el(0x28, 0x103, b"US", struct.pack("<H", 0)) \
        + el(0x7FE0, 0x10, b"OW", px.tobytes())
    p = tmp_path / "t.dcm"
    p.write_bytes(b"\0" * 128 + b"DICM" + meta + body)
    arr, info = dicom_raw.read_pixels(str(p))
    assert np.array_equal(arr, px) and info["BitsStored"] == 12 and info["Rows"] == 6
    bad = tmp_path / "c.dcm"
    bad.write_bytes(b"\0" * 128 + b"DICM" + el(2, 0x10, b"UI", b"1.2.840.10008.1.2.4.80") + body)
    with pytest.raises(ValueError):
        dicom_raw.read_pixels(str(bad))


@pytest.mark.gpu
def test_metrics_take_dcm_paths(tmp_path):
    """File-path inputs of AnalisadorMSE (src/mse.py:82-84): the range comes from BitsStored, not from the data."""
    from codec_tcc_b200 import mse as M

    def el(g, e, vr, val):
        if vr in (b"OB", b"OW"):
            return struct.pack("<HH2sHI", g, e, vr, 0, len(val)) + val
        return struct.pack("<HH2sH", g, e, vr, len(val)) + val

    def write(path, px):
        body = el(0x28, 2, b"US", struct.pack("<H", 1)) + el(0x28, 0x10, b"US", struct.pack("<H", px.shape[0])) \
            + el(0x28, 0x11, b"US", struct.pack("<H", px.shape[1])) + el(0x28, 0x100, b"US", struct.pack("<H", 16)) \
            + el(0x28, 0x101, b"US", struct.pack("<H", 12)) + el(0x7FE0, 0x10, b"OW", px.astype("<u2").tobytes())
        open(path, "wb").write(b"\0" * 128 + b"DICM" + el(2, 0x10, b"UI", b"1.2.840.10008.1.2.1\0") + body)

    rng = np.random.default_rng(5)
    a = rng.integers(0, 3000, (40, 64)).astype(np.uint16)
    b = a.copy(); b[3, 5] += 7; b[10, 1] -= 2
    pa, pb = str(tmp_path / "a.dcm"), str(tmp_path / "b.dcm")
    write(pa, a); write(pb, b)
    an = M.AnalisadorMSE()
    mse, top = an.calcular_mse(pa, pb)
    assert float(top) == 4095.0 and float(mse) == (49 + 4) / a.size      # both ranges 2**12-1: no rescaling
    mse2, top2 = an.calcular_mse(a, b)                                     # arrays: ranges are the maxima
    assert float(top2) == float(max(a.max(), b.max()))
    f64, maxv, bits = an.carregar_imagem(pa)
    assert maxv == 4095 and bits == 12 and np.array_equal(f64, a)
