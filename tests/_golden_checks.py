"""Shared checkers: the same assertions run against the CPU oracle (not gpu)
and against the CUDA product (gpu).  ``impl`` is any object exposing the
reference's function names (src/codec.py) and ``metrics`` one exposing
``AnalisadorMSE``'s methods (src/mse.py)."""
from __future__ import annotations

import hashlib

import numpy as np

MAIN_MESSAGE = "Mensagem de teste para esteganografia!"  # src/codec.py:863


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rand_bitstring(n, seed):
    rng = np.random.default_rng(seed)
    return "".join("1" if b else "0" for b in rng.integers(0, 2, n).tolist())


def case_bits(case, img, impl):
    """Rebuild the payload make_golden.py used for a given lsb case."""
    npx = img.size
    if case["n_bits"] == 8 * len(MAIN_MESSAGE):
        return impl.message_to_bits(MAIN_MESSAGE)
    if case["n_bits"] == min(60000, npx // 2):
        return rand_bitstring(case["n_bits"], 5)
    assert case["n_bits"] == 4 * npx
    return rand_bitstring(case["n_bits"], 6)


def check_lsb_case(impl, metrics, img, case, quiet=None):
    """One row of golden['images'][name]['lsb_cases'] -- bit-exact on every
    array (sha256), exact on the integers, <=1e-9 relative on float scalars."""
    bits = case_bits(case, img, impl)
    g, l = impl.adaptive_modalities_decomposition(img, beta=case["beta"])
    s = len(l)
    assert s == case["s"]
    if case["embedder"] == "hybrid":
        sp, bm, used, lens, idx = impl.lsb_embed_block_then_multiplane(
            l, bits, search_block_size=case["sbs"], align_across_planes=case["align"])
    else:
        sp, bm, used, lens, idx = impl.lsb_embed_multi_plane(l, bits)
    assert int(used) == case["total_used"]
    assert [int(v) for v in lens] == case["segments_lengths"]
    assert [int(v) for v in idx] == case["segment_indices"]
    assert all(b.dtype == np.uint8 and b.shape == img.shape for b in bm)
    assert all(p.dtype == img.dtype for p in sp)
    assert sha(np.stack(sp)) == case["planes_sha"]
    assert sha(np.stack(bm)) == case["bitmaps_sha"]
    stego = impl.merge_modalities(g, sp)
    assert str(stego.dtype) == case["stego_dtype"]
    assert sha(stego) == case["stego_sha"]
    meta = {"s": s, "segments_indices": idx, "segments_lengths": lens}
    decoded = impl.decode_message(impl.extract_local_planes(stego, s), [b.ravel() for b in bm], meta)
    assert len(decoded) == case["decoded_len"]
    assert hashlib.sha256(decoded.encode("utf-8")).hexdigest() == case["decoded_sha"]
    # N4 (true inverse, not in the reference): on the reference's own stego image (sha pinned above)
    if hasattr(impl, "recover_cover"):
        sub = stego if stego.dtype == img.dtype else stego.astype(img.dtype)
        assert np.array_equal(impl.recover_cover(sub, bm), img)
        meta4 = {"s": s, "segments_indices": idx, "segments_lengths": lens, "hybrid": case["embedder"] == "hybrid",
                 "align_across_planes": case.get("align", False), "message_bits": len(bits) if case["embedder"] == "hybrid" else None,
                 "start_offset": impl.hybrid_start_offset(l[0], case["sbs"]) if case["embedder"] == "hybrid" else 0}
        got = impl.extract_message_bits(sub, meta4)
        if int(used) == len(bits):
            assert got == bits
        else:  # truncated segments: every embedded bit is still read back from where it was written
            assert len(got) == int(used)
    # XOR side information really restores the cover
    back = impl.merge_modalities(g, [p ^ b.astype(p.dtype) for p, b in zip(sp, bm)])
    assert np.array_equal(back, img if back.dtype == img.dtype else img.astype(back.dtype))
    if metrics is not None:
        m, rng_ = metrics.calcular_mse(img, stego)
        if int(img.max()) == int(stego.max()):
            assert float(m) == case["mse"], (float(m), case["mse"])  # exact: int64 SSE / N
        else:  # the reference rescales both images in floating point (src/mse.py:101-106)
            assert abs(float(m) - case["mse"]) <= 1e-9 * case["mse"], (float(m), case["mse"])
        assert float(rng_) == case["max_range"]
        psnr = metrics.calcular_psnr(m, rng_)
        assert abs(float(psnr) - case["psnr"]) <= 1e-9 * abs(case["psnr"])
        ssim = metrics.calcular_ssim_simples(img, stego)
        assert abs(float(ssim) - case["ssim"]) <= 1e-9
        if hasattr(metrics, "difference_stats"):
            mean_abs, max_abs, changed, pct = metrics.difference_stats(img, stego)
            assert int(changed) == case["px_changed"]
            assert float(max_abs) == case["max_abs"]
            assert abs(float(mean_abs) - case["mean_abs"]) <= 1e-12 * max(1.0, case["mean_abs"])


def check_entropy_and_split(impl, img, rec):
    assert sha(img) == rec["sha"]
    assert float(impl.calculate_entropy(img)) == rec["entropy"]
    for beta, s in rec["split"].items():
        g, l = impl.adaptive_modalities_decomposition(img, beta=float(beta))
        assert len(l) == s and len(g) == 8 * img.dtype.itemsize - s
        for i, p in enumerate(l):
            assert p.dtype == img.dtype and np.array_equal(p, (img >> i) & 1)
        for i, p in enumerate(g):
            assert np.array_equal(p, (img >> (i + s)) & 1)
    for i, mi in enumerate(rec["mi"]):
        got = impl.calculate_mutual_information((img >> i) & 1, img)
        assert float(got) == mi, (i, float(got), mi)


# ---- inputs of the golden entries added in round 2 (shared with tests/golden/make_golden.py) ----
def general_mi_cases():
    """(name, plane, image): planes that are NOT bit planes of the image -- the general branch of
    calculate_mutual_information (src/codec.py:504-559 accepts any non-negative integer plane)."""
    from codec_tcc_b200.synth import synth_image

    a, b = synth_image(129, 70, 255, 13), synth_image(129, 70, 255, 31)
    c, d = synth_image(120, 96, 4095, 12), synth_image(120, 96, 4095, 33)
    rng = np.random.default_rng(77)
    return [
        ("u8_other_image_bit3", (b >> 3) & 1, a),
        ("u8_random_bits", rng.integers(0, 2, a.shape).astype(np.uint8), a),
        ("u8_three_values", ((b >> 6) % 3).astype(np.uint8), a),
        ("u16_other_image_bit5", ((d >> 5) & 1).astype(np.uint16), c),
        ("u16_threshold_of_image", (c > 2000).astype(np.uint16), c),
        ("u16_image_u8_plane", (d >> 11).astype(np.uint8), c),
    ]


def float_metric_cases():
    """(name, img1, img2): inputs the reference handles through np.array(img, dtype=np.float64) (src/mse.py:85)
    that are not integer-valued 8/16-bit pixel data."""
    from codec_tcc_b200.synth import synth_image

    a = synth_image(120, 90, 4095, 21).astype(np.float64)
    b = a + np.random.default_rng(9).normal(0.0, 2.5, a.shape)
    return [
        ("fractional", a, b),
        ("negative", a - 1000.25, b - 1000.0),
        ("different_ranges", a * 0.5, b),
        ("wide_integers", (a * 40).astype(np.int64), (b * 40).astype(np.int64)),
    ]
