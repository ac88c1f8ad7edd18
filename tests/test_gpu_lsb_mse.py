"""GPU parity: the reference's bit-plane LSB path and distortion metrics
(rows a1-a9), CUDA through the C ABI, against the golden vectors produced by
the unmodified reference and against the numpy restatement on random inputs."""
import numpy as np
import pytest

from codec_tcc_b200 import codec, mse
from codec_tcc_b200.synth import synth_image, synth_saturated

from oracle import codec_numpy as OC
from oracle import mse_numpy as OM

import _golden_checks as GC

pytestmark = pytest.mark.gpu

IMAGES = ["pe", "torax", "synth16_257x301", "synth12_300x200", "synth8_129x70", "sat12_96x160"]


class _Metrics(mse.AnalisadorMSE):
    pass


@pytest.mark.parametrize("name", IMAGES)
def test_entropy_mi_split(golden, golden_images, name):
    GC.check_entropy_and_split(codec, golden_images[name], golden["images"][name])


@pytest.mark.parametrize("name", IMAGES)
@pytest.mark.parametrize("case_idx", range(6))
def test_lsb_golden_cases(golden, golden_images, name, case_idx):
    GC.check_lsb_case(codec, _Metrics(), golden_images[name], golden["images"][name]["lsb_cases"][case_idx])


def _bits(n, seed):
    rng = np.random.default_rng(seed)
    return "".join("1" if b else "0" for b in rng.integers(0, 2, n).tolist())


@pytest.mark.parametrize("idx", range(7))
def test_lsb_random_vs_restatement(idx):
    img, beta, sbs = [
        (synth_image(67, 45, 255, 1), 0.5, 8),
        (synth_image(50, 130, 4095, 2), 0.7, 16),
        (synth_image(33, 33, 65535, 3), 0.9, 7),
        (synth_saturated(40, 56, 255, 4), 0.3, 4),
        (np.zeros((20, 24), np.uint16), 0.8, 8),
        (np.full((9, 9), 255, np.uint8), 0.8, 16),
        (synth_image(400, 1000, 4095, 8), 0.8, 16),
    ][idx]
    g0, l0 = OC.adaptive_modalities_decomposition(img, beta=beta)
    g1, l1 = codec.adaptive_modalities_decomposition(img, beta=beta)
    assert len(l0) == len(l1) and all(np.array_equal(a, b) and a.dtype == b.dtype for a, b in zip(l0 + g0, l1 + g1))
    for n_payload in (0, 1, 5, 1000, img.size * 3):
        bits = _bits(n_payload, idx)
        for align in (False, True):
            r = OC.lsb_embed_block_then_multiplane(l0, bits, search_block_size=sbs, align_across_planes=align)
            o = codec.lsb_embed_block_then_multiplane(l1, bits, search_block_size=sbs, align_across_planes=align)
            _same(r, o)
        r = OC.lsb_embed_multi_plane(l0, bits)
        o = codec.lsb_embed_multi_plane(l1, bits)
        _same(r, o)
        meta = {"s": len(l0), "segments_indices": r[4], "segments_lengths": r[3]}
        assert OC.decode_message(r[0], [b.ravel() for b in r[1]], meta) == \
            codec.decode_message(o[0], [b.ravel() for b in o[1]], meta)
        st0, st1 = OC.merge_modalities(g0, r[0]), codec.merge_modalities(g1, o[0])
        assert st0.dtype == st1.dtype and np.array_equal(st0, st1)
        for a, b in zip(OC.extract_local_planes(st0, len(l0)), codec.extract_local_planes(st1, len(l0))):
            assert a.dtype == b.dtype and np.array_equal(a, b)


def _same(r, o):
    assert r[2] == o[2] and list(r[3]) == list(o[3]) and list(r[4]) == list(o[4])
    for a, b in zip(r[0], o[0]):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    for a, b in zip(r[1], o[1]):
        assert a.dtype == b.dtype == np.uint8 and np.array_equal(a, b)


@pytest.mark.parametrize("shape,hi", [((257, 301), 65536), ((64, 64), 65536), ((1000, 1003), 65536), ((300, 331), 20000),
                                      ((129, 77), 4096), ((1, 5), 65536), ((2100, 2050), 65536)])
def test_histogram_full_range_16bit(shape, hi, monkeypatch):
    """Row a5's histogram on 16-bit data that leaves the 16 K-bin window (packed full-range counters) and on
    data inside it (window kernel), forced both ways: exact counts and plane population counts."""
    from codec_tcc_b200 import codec
    rng = np.random.default_rng(shape[0] * 7 + hi)
    img = rng.integers(0, hi, shape, dtype=np.uint16)
    img[0, 0] = hi - 1
    flat = np.concatenate([np.zeros(1, np.uint16), img.reshape(-1)])[1:].reshape(shape)  # a view that is only 2-byte aligned
    ref = np.bincount(img.reshape(-1), minlength=65536)
    ones = np.array([int(((img >> b) & 1).sum()) for b in range(16)])
    for force_window in (False, True):
        if force_window:
            monkeypatch.setenv("PEEB_HIST_WINDOW_ONLY", "1")
        for a in (img, flat):
            hist, po = codec._image_histogram(a)
            assert np.array_equal(hist, ref) and np.array_equal(po, ones)
    # many hits on one bin: more than a 16-bit counter holds
    big = np.full((700, 700), 60001, np.uint16)
    big[::7, ::3] = 60000
    hist, _ = codec._image_histogram(big)
    assert np.array_equal(hist, np.bincount(big.reshape(-1), minlength=65536))


def test_metrics_scalars(golden):
    an = mse.AnalisadorMSE()
    sc = golden["scalars"]
    m, r = an.calcular_mse([[10, 20], [30, 40]], [[10, 20], [30, 41]])
    assert abs(float(m) - sc["mse_norm_small"][0]) <= 1e-12 and float(r) == 41.0
    a = synth_image(120, 90, 4095, 21)
    b = a.copy(); b[5, 7] += 900; b[60:70, 10:50] ^= 3
    m, r = an.calcular_mse(a, b)
    assert abs(float(m) - sc["mse_norm_synth"][0]) <= 1e-9 * sc["mse_norm_synth"][0] and float(r) == sc["mse_norm_synth"][1]
    assert abs(float(an.calcular_ssim_simples(a, b)) - sc["ssim_norm_synth"]) <= 1e-9
    m, r = an.calcular_mse(a, a)
    assert [float(m), float(r)] == sc["mse_same"] and an.calcular_psnr(m, r) == float("inf")
    assert float(an.calcular_ssim_simples(a, a)) == 1.0
    with pytest.raises(ValueError):
        an.calcular_mse(np.zeros((2, 3), np.uint8), np.zeros((3, 2), np.uint8))


@pytest.mark.parametrize("shape", [(1, 1), (7, 13), (512, 512), (1000, 1003), (3000, 3000)])
@pytest.mark.parametrize("maxval", [255, 4095, 65535])
def test_moments_exact(shape, maxval):
    """int64 moments against numpy; MSE exact, PSNR/SSIM within 1e-9 relative
    of the float64 element-wise restatement (BASELINE.json tolerance)."""
    a = synth_image(shape[0], shape[1], maxval, 5)
    rng = np.random.default_rng(shape[0])
    b = a.copy()
    k = max(1, a.size // 7)
    ys, xs = rng.integers(0, shape[0], k), rng.integers(0, shape[1], k)
    b[ys, xs] = rng.integers(0, maxval + 1, k).astype(a.dtype)
    b.flat[0] = a.max()  # keep equal maxima: the exact (non-normalised) branch
    if b.max() != a.max():
        b[b > a.max()] = a.max()
    mm = mse.image_moments(a, b)
    ai, bi = a.astype(np.int64), b.astype(np.int64)
    d = ai - bi
    assert mm["sse"] == int((d * d).sum()) and mm["sad"] == int(np.abs(d).sum())
    assert mm["max_abs"] == int(np.abs(d).max()) and mm["changed"] == int((d != 0).sum())
    assert mm["sum_a"] == int(ai.sum()) and mm["sum_b"] == int(bi.sum()) and mm["sum_ab"] == int((ai * bi).sum())
    assert mm["sum_aa"] == int((ai * ai).sum()) and mm["sum_bb"] == int((bi * bi).sum())
    assert mm["max_a"] == int(a.max()) and mm["max_b"] == int(b.max()) and mm["n"] == a.size
    an = mse.AnalisadorMSE()
    m1, r1 = an.calcular_mse(a, b)
    m0, r0 = OM.calcular_mse(a, b)
    assert float(m1) == float(m0) and float(r1) == float(r0)  # bit-exact MSE
    p1, p0 = an.calcular_psnr(m1, r1), OM.calcular_psnr(m0, r0)
    assert p1 == p0 or abs(p1 - p0) <= 1e-9 * abs(p0)
    s1, s0 = float(an.calcular_ssim_simples(a, b)), float(OM.calcular_ssim_simples(a, b))
    assert abs(s1 - s0) <= 1e-9
    st1, st0 = an.difference_stats(a, b), OM.difference_stats(a, b)
    assert float(st1[0]) == float(st0[0]) and float(st1[1]) == float(st0[1]) and int(st1[2]) == int(st0[2])
    assert abs(float(st1[3]) - float(st0[3])) <= 1e-12


def test_normalisation_branch_and_mixed_inputs():
    an = mse.AnalisadorMSE()
    a = synth_image(200, 300, 4095, 2)
    b = a.copy(); b[0, 0] = 4095 if a.max() < 4095 else a.max() - 5; b[50:60] ^= 1
    m1, r1 = an.calcular_mse(a, b)
    m0, r0 = OM.calcular_mse(a, b)
    assert abs(float(m1) - float(m0)) <= 1e-9 * float(m0) and float(r1) == float(r0)
    assert abs(float(an.calcular_ssim_simples(a, b)) - float(OM.calcular_ssim_simples(a, b))) <= 1e-9
    # list / int64 / uint8-vs-uint16 inputs
    m1, _ = an.calcular_mse(a.astype(np.int64), b.tolist())
    assert abs(float(m1) - float(m0)) <= 1e-9 * float(m0)
    c = synth_image(64, 64, 255, 3)
    m1, r1 = an.calcular_mse(c, c.astype(np.uint16) + 1)
    m0, r0 = OM.calcular_mse(c, c.astype(np.uint16) + 1)
    assert abs(float(m1) - float(m0)) <= 1e-9 * float(m0) and float(r1) == float(r0)
    res = an.analisar_par_arrays(a, b, "par")
    assert set(res) >= {"mse", "psnr", "ssim", "diferenca_media", "diferenca_max", "percentual_mudanca"}


def test_float_route_of_the_metrics(golden):
    """Inputs that are not integer-valued 8/16-bit data take the float64 kernels (the reference's own
    np.array(img, dtype=np.float64) route, src/mse.py:85): golden values of the unmodified reference, 1e-9."""
    an = mse.AnalisadorMSE()
    for name, x, y in GC.float_metric_cases():
        want = golden["scalars"]["float_metrics"][name]
        m, r = an.calcular_mse(x, y)
        assert abs(float(m) - want["mse"]) <= 1e-9 * want["mse"], (name, float(m), want["mse"])
        assert abs(float(r) - want["max_range"]) <= 1e-12 * abs(want["max_range"]), name
        assert abs(float(an.calcular_ssim_simples(x, y)) - want["ssim"]) <= 1e-9, name
        st = an.difference_stats(x, y)
        assert abs(float(st[0]) - want["mean_abs"]) <= 1e-9 * want["mean_abs"] and int(st[2]) == want["changed"], name
        assert abs(float(st[1]) - want["max_abs"]) <= 1e-12 * want["max_abs"], name
        m0, r0 = OM.calcular_mse(x, y)
        assert abs(float(m) - float(m0)) <= 1e-9 * float(m0)
    # lists of floats, one float and one integer image
    a = synth_image(64, 80, 4095, 2)
    m1, r1 = an.calcular_mse((a + 0.25).tolist(), a)   # (the maxima differ: the range normalisation applies)
    m0, r0 = OM.calcular_mse((a + 0.25).tolist(), a)
    assert abs(float(m1) - float(m0)) <= 1e-9 * float(m0) and abs(float(r1) - float(r0)) <= 1e-12 * float(r0)
    with pytest.raises(ValueError):
        an.calcular_mse(np.zeros((2, 3)) + 0.5, np.zeros((3, 2)))


def test_mutual_information_of_arbitrary_planes(golden):
    """calculate_mutual_information for planes that are not bit planes of the image (src/codec.py:504-559 accepts
    any): bit for bit the reference's value (golden) and the restatement's."""
    for name, plane, img in GC.general_mi_cases():
        got = float(codec.calculate_mutual_information(plane, img))
        assert got == golden["scalars"]["mi_general"][name], (name, got)
        assert got == float(OC.calculate_mutual_information(plane, img))
    img = synth_image(64, 64, 255, 3)
    assert codec.calculate_mutual_information(np.zeros_like(img), img) == 0.0          # constant plane
    assert codec.calculate_mutual_information(img & 1, np.full_like(img, 7)) == 0.0    # constant image
    with pytest.raises(ValueError):
        codec.calculate_mutual_information(np.zeros((3, 3), np.uint8), img)


@pytest.mark.parametrize("name", IMAGES)
def test_embed_pipeline_equals_chained_reference_flow(golden, golden_images, name):
    """The device-resident encode flow (src/codec.py:868-880 in one call) against the golden
    vectors of the chained reference functions."""
    img = golden_images[name]
    for case in golden["images"][name]["lsb_cases"]:
        bits = GC.case_bits(case, img, codec)
        stego, bitmaps, meta = codec.embed_pipeline(img, bits, beta=case["beta"], search_block_size=case["sbs"],
                                                    align_across_planes=case["align"],
                                                    hybrid=case["embedder"] == "hybrid")
        assert meta["s"] == case["s"] and meta["total_used"] == case["total_used"]
        assert [int(v) for v in meta["segments_lengths"]] == case["segments_lengths"]
        assert [int(v) for v in meta["segments_indices"]] == case["segment_indices"]
        assert str(stego.dtype) == case["stego_dtype"] and GC.sha(stego) == case["stego_sha"]
        assert bitmaps.dtype == np.uint8 and GC.sha(bitmaps) == case["bitmaps_sha"]


@pytest.mark.parametrize("idx", range(5))
def test_true_inverse_vs_restatement(idx):
    """N4: recover_cover / extract_message_bits on random shapes, wrapping segments,
    truncated segments and unaligned sizes, against the numpy restatement and the inputs."""
    img, beta, sbs = [
        (synth_image(67, 45, 255, 11), 0.5, 8),
        (synth_image(50, 130, 4095, 12), 0.7, 16),
        (synth_image(33, 33, 65535, 13), 0.9, 7),
        (synth_saturated(40, 56, 255, 14), 0.3, 4),
        (synth_image(256, 512, 4095, 15), 0.8, 16),
    ][idx]
    g, l = codec.adaptive_modalities_decomposition(img, beta=beta)
    s = len(l)
    for n_payload in (0, 1, 7, 1000, img.size, img.size * 3):
        bits = _bits(n_payload, 40 + idx)
        for hybrid, align in ((True, False), (True, True), (False, False)):
            if hybrid:
                sp, bm, used, lens, order = codec.lsb_embed_block_then_multiplane(l, bits, search_block_size=sbs, align_across_planes=align)
            else:
                sp, bm, used, lens, order = codec.lsb_embed_multi_plane(l, bits)
            stego = codec.merge_modalities(g, sp).astype(img.dtype)
            meta = {"s": s, "segments_indices": order, "segments_lengths": lens, "hybrid": hybrid, "align_across_planes": align,
                    "start_offset": codec.hybrid_start_offset(l[0], sbs) if hybrid else 0,
                    "message_bits": len(bits) if hybrid else None}
            assert np.array_equal(codec.recover_cover(stego, bm), img)
            assert np.array_equal(OC.recover_cover(stego, bm), img)
            got = codec.extract_message_bits(stego, meta)
            assert got == OC.extract_message_bits(stego, meta)
            assert len(got) == used
            # the exact inverse of the embedder: its segments (src/codec.py:267-272), as far as they fitted
            assert got == "".join(seg[:img.size] for seg in codec.distribute_message_segments(l, bits)[0])
            if used == len(bits) and len(bits) >= 4 * s:
                assert got == bits
    msg = "ol\u00e1 B200"
    sp, bm, used, lens, order = codec.lsb_embed_multi_plane(l, codec.message_to_bits(msg))
    if used == 8 * len(msg):  # message_to_bits is one byte per character (src/codec.py:239-240): latin-1 text only
        text = codec.extract_message(codec.merge_modalities(g, sp).astype(img.dtype), {"s": s, "segments_indices": order, "segments_lengths": lens})
        assert text.encode("utf-8", errors="replace") is not None


def test_embed_pipeline_with_device_coded_bitmaps(golden_images):
    """bitmaps_as="pbr": the blob equals container.pack_bitmaps(bitmaps, coding="pbr") of the plain call and decodes
    to the same arrays (N2: the bitmaps never cross PCIe as one byte per pixel)."""
    from codec_tcc_b200 import container
    from oracle import bitcode_numpy as BN
    for name in ("pe", "synth12_300x200"):
        img = golden_images[name]
        bits = codec.message_to_bits("Mensagem de teste para esteganografia!")
        stego0, bm0, meta0 = codec.embed_pipeline(img, bits, beta=0.4, search_block_size=16)
        stego1, blob, meta1 = codec.embed_pipeline(img, bits, beta=0.4, search_block_size=16, bitmaps_as="pbr")
        assert np.array_equal(stego0, stego1) and meta0 == meta1
        assert blob == BN.encode(bm0)
        back = container.unpack_bitmaps(blob, meta0["s"])
        assert all(np.array_equal(b, m.ravel()) for b, m in zip(back, bm0))
        assert len(blob) < bm0.size // 64


def test_tile_search_on_bright_16_bit_planes_with_large_tiles():
    """n * sum(v^2) beyond 2^53 (16-bit values, 256x256 tiles): the tile search still picks the reference's tile (exact
    scores for the shortlist instead of a float64 difference that cancels)."""
    rng = np.random.default_rng(17)
    for seed in range(3):
        base = rng.integers(60000, 65536, (700, 900)).astype(np.uint16)
        base[256:512, 256:512] -= rng.integers(0, 3, (256, 256)).astype(np.uint16) * 20000   # the widest spread
        base[0:256, 512:768] -= rng.integers(0, 3, (256, 256)).astype(np.uint16) * 19999      # a close second
        assert codec.hybrid_start_offset(base, 256) == OC.best_tile_offset(base, 256)
        assert codec.hybrid_start_offset(base, 128) == OC.best_tile_offset(base, 128)


@pytest.mark.parametrize("seed", range(8))
def test_segment_read_back_arbitrary_segments(seed):
    """peeb_lsb_extract with segments the embedders never produce together: starts at every alignment, segments
    that wrap around the end of the image, lengths that end inside a 1024-bit warp item, image sizes that are not a
    multiple of 8, 8- and 16-bit pixels -- the aligned-block form and the generic form of the kernel against numpy."""
    from codec_tcc_b200 import _cabi

    rng = np.random.default_rng(100 + seed)
    dtype = np.uint16 if seed % 2 else np.uint8
    s = int(rng.integers(1, 8 * dtype().itemsize + 1))
    n = int(rng.integers(3000, 150000))
    img = rng.integers(0, np.iinfo(dtype).max + 1, n).astype(dtype)
    start = rng.integers(0, n, s).astype(np.int64)
    length = rng.integers(0, n + 1, s).astype(np.int64)
    length[rng.integers(0, s)] = n                      # one segment goes all the way round
    if s > 2:
        length[rng.integers(0, s)] = 0
    order = rng.permutation(s)
    off = np.zeros(s, np.int64)
    at = 0
    for p in order:
        off[p] = at
        at += int(length[p])
    total = at
    out = np.zeros((total + 7) // 8 + 8, np.uint8)
    ws = _cabi.workspace(None)
    _cabi.check(_cabi.lib().peeb_lsb_extract_h(ws.handle, _cabi.ptr(img), n, img.dtype.itemsize, s, _cabi.ptr(start),
                                               _cabi.ptr(length), _cabi.ptr(off), total, _cabi.ptr(out)), "peeb_lsb_extract_h")
    want = np.zeros(total, np.uint8)
    for p in range(s):
        seg = np.roll(img, -int(start[p]))[:int(length[p])]
        want[off[p]:off[p] + length[p]] = (seg >> p) & 1
    assert np.array_equal(np.unpackbits(out)[:total], want)
    assert not out[(total + 7) // 8:].any()
