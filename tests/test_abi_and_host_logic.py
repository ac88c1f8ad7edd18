"""CPU: the C-ABI library loads and exports every symbol include/peeb200.h
declares (no compute without a GPU), and the host-side logic of the product
package (everything above the C ABI that is not per-pixel work)."""
import os
import re

import numpy as np
import pytest

from codec_tcc_b200 import _cabi, codec, mse, pee
from codec_tcc_b200.synth import synth_image

from oracle import codec_numpy as OC
from oracle import mse_numpy as OM
from oracle import pee_numpy as PN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "peeb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(peeb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _cabi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/peeb200.h but not exported"
    assert sorted(_cabi.SIGNATURES) == declared, "ctypes table and header disagree"
    assert L.peeb_abi_version() == 1
    assert _cabi.payload_bytes(0) == 8 and _cabi.payload_bytes(1) == 12 and _cabi.payload_bytes(33) == 16


def test_compute_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.PeebError):
        mse.AnalisadorMSE().calcular_mse(np.zeros((4, 4), np.uint8), np.zeros((4, 4), np.uint8))
    with pytest.raises(_cabi.PeebError):
        pee.pee_embed(np.zeros((8, 8), np.uint8), b"", 1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "codec_tcc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("test oracle", ""), f"{f} mentions the oracle"


def test_segment_plan_matches_golden(golden):
    for key, rec in golden["segments"].items():
        s, total = (int(v) for v in key.split(":"))
        segs, sizes, order = codec.distribute_message_segments([None] * s, "0" * total)
        assert sizes == rec["sizes"] and order == rec["order"] and [len(x) for x in segs] == rec["seg_lens"]
    assert codec.message_to_bits("Olá, DICOM ✓") == golden["scalars"]["message_bits_hex"]


def test_plane_information_from_histogram(golden, golden_images):
    """The histogram-side entropy code of the product against the golden MI
    values (device histogram replaced by np.bincount here)."""
    for name in ("pe", "torax", "synth12_300x200"):
        img = golden_images[name]
        nbins = 256 if img.dtype == np.uint8 else 65536
        hist = np.bincount(img.ravel(), minlength=nbins).astype(np.int64)
        assert float(codec._entropy_from_counts(hist, img.size)) == golden["images"][name]["entropy"]
        for bit, want in enumerate(golden["images"][name]["mi"]):
            ones = int(np.count_nonzero((img >> bit) & 1))
            assert float(codec._plane_information(hist, ones, img.size, bit, nbins)) == want


def _np_moments(a, b):
    a = a.astype(np.int64).ravel(); b = b.astype(np.int64).ravel()
    d = a - b
    return {"sse": int((d * d).sum()), "sad": int(np.abs(d).sum()), "max_abs": int(np.abs(d).max()),
            "changed": int((d != 0).sum()), "sum_a": int(a.sum()), "sum_b": int(b.sum()),
            "sum_aa": int((a * a).sum()), "sum_bb": int((b * b).sum()), "sum_ab": int((a * b).sum()),
            "max_a": int(a.max()), "max_b": int(b.max()), "n": int(a.size)}


def test_metrics_from_moments_match_reference_arithmetic(golden):
    A = mse.AnalisadorMSE
    a = synth_image(120, 90, 4095, 21)
    b = a.copy(); b[5, 7] += 900; b[60:70, 10:50] ^= 3
    m, r = A._mse_from(_np_moments(a, b))
    want = golden["scalars"]["mse_norm_synth"]
    assert abs(float(m) - want[0]) <= 1e-12 * want[0] and float(r) == want[1]
    assert abs(float(A._ssim_from(_np_moments(a, b))) - golden["scalars"]["ssim_norm_synth"]) <= 1e-12
    m, r = A._mse_from(_np_moments(np.array([[10, 20], [30, 40]]), np.array([[10, 20], [30, 41]])))
    assert abs(float(m) - 0.21875) <= 1e-15 and float(r) == 41.0
    rng = np.random.default_rng(3)
    for maxval in (255, 4095, 65535):
        x = synth_image(50, 70, maxval, 4)
        y = x.copy(); y[rng.integers(0, 50, 40), rng.integers(0, 70, 40)] ^= 1
        mm = _np_moments(x, y)
        m0, r0 = OM.calcular_mse(x, y)
        m1, r1 = A._mse_from(mm)
        assert float(m1) == float(m0) and float(r0) == float(r1) or abs(float(m1) - float(m0)) <= 1e-12 * float(m0)
        assert abs(float(A._ssim_from(mm)) - float(OM.calcular_ssim_simples(x, y))) <= 1e-12
    an = mse.AnalisadorMSE()
    assert an.calcular_psnr(0.0, 4095) == float("inf")
    assert float(an.calcular_psnr(1.0)) == golden["scalars"]["psnr_default"]
    assert float(an.calcular_psnr(0.37, 4095)) == golden["scalars"]["psnr_4095"]


def _np_tile_moments(plane, sbs):
    h, w = plane.shape
    out = []
    for y in range(0, h, sbs):
        for x in range(0, w, sbs):
            t = plane[y:y + sbs, x:x + sbs].astype(np.int64)
            out.append((int(t.sum()), int((t * t).sum())))
    return np.array(out, np.int64)


@pytest.mark.parametrize("shape,sbs", [((64, 64), 16), ((257, 301), 16), ((129, 70), 8), ((45, 33), 7), ((9, 9), 16)])
def test_tile_argmax_logic(shape, sbs):
    for seed in range(3):
        img = synth_image(shape[0], shape[1], 4095, seed)
        for bit in (0, 3):
            plane = (img >> bit) & 1
            got = codec._tile_argmax_from_moments(plane, sbs, _np_tile_moments(plane, sbs))
            assert got == OC.best_tile_offset(plane, sbs)
    flat = np.zeros(shape, np.uint16)
    assert codec._tile_argmax_from_moments(flat, sbs, _np_tile_moments(flat, sbs)) == 0


def test_payload_packing_and_threshold_estimate():
    packed, n = pee.pack_payload("1011001110")
    assert n == 10 and packed.tolist() == [0b10110011, 0b10000000]
    packed, n = pee.pack_payload(b"\xff\x00", 12)
    assert n == 12 and packed.tolist() == [255, 0]
    with pytest.raises(ValueError):
        pee.pack_payload("10a")
    with pytest.raises(ValueError):
        pee.pack_payload(b"\x00", 9)
    img = synth_image(60, 50, 4095, 2)
    hist = PN.error_histogram(img, 12)
    for nb in (0, 10, 500, 2000, 10 ** 7):
        assert pee.estimate_threshold(hist, nb) == PN.estimate_T(hist, nb)


def test_result_arrays_come_from_the_pool_or_plain_numpy():
    """_cabi.out_empty: small results are ordinary arrays; large ones are page-locked blocks from a pool when a device is
    there, ordinary arrays when the host will not pin memory (no GPU here) -- either way a writable array of the shape."""
    a = _cabi.out_empty((3, 5), np.uint16)
    assert a.shape == (3, 5) and a.dtype == np.uint16 and a.flags.writeable
    b = _cabi.out_empty((3000, 3000), np.uint8)
    assert b.shape == (3000, 3000) and b.dtype == np.uint8 and b.flags.c_contiguous
    b[...] = 7
    assert int(b.sum()) == 7 * 9_000_000
    del b
    c = _cabi.out_empty((3000, 3000), np.uint8)   # a released block may be handed out again
    c[0, 0] = 1
    assert c[0, 0] == 1
