"""GPU parity: the CUDA PEE path (through the C ABI) against the CPU oracles,
bit for bit -- marked image, location map, info, extracted payload, recovered
image.  PEE parity is UNPINNED (no PEE in the reference): the
oracle follows SURVEY.md Appendix A."""
import numpy as np
import pytest

from codec_tcc_b200 import _cabi, pee
from codec_tcc_b200.synth import random_payload, synth_batch, synth_image, synth_saturated

from oracle import pee_c as PC
from oracle import pee_numpy as PN

pytestmark = pytest.mark.gpu

SHAPES = [(3, 3), (3, 40), (40, 3), (5, 7), (17, 129), (64, 64), (70, 131), (257, 301), (64, 512), (100, 1040),
          (33, 2048)]


def _cap(img, T, bd):
    _, _, info = PC.embed(img, np.zeros(img.size // 8 + 8, np.uint8), 0, T, bd)
    return info["capacity"]


def _check_one(img, T, bd, frac, seed):
    n_bits = int(_cap(img, T, bd) * frac)
    pay = random_payload(n_bits, seed)
    m0, lm0, i0 = PC.embed(img, pay, n_bits, T, bd)
    assert i0.pop("status") == 0
    m1, lm1, i1 = pee.pee_embed(img, pay, T, bd, n_bits=n_bits)
    assert i1 == i0, (i1, i0)
    assert np.array_equal(lm1, lm0), "location map differs"
    assert np.array_equal(m1, m0), f"marked differs at {np.argwhere(m1 != m0)[:5]}"
    p1, r1 = pee.pee_extract(m1, lm1, T, n_bits, bd)
    assert np.array_equal(r1, img), f"recovered differs at {np.argwhere(r1 != img)[:5]}"
    assert np.array_equal(p1, pay), "payload differs"


@pytest.fixture(params=["bands", "cluster"])
def kernel_path(request, monkeypatch):
    """Both families of PEE kernels: the band kernels (batches) and the cluster path for small images (one thread-block
    cluster per image); PEEB_CLUSTER forces the choice where an image qualifies for both."""
    monkeypatch.setenv("PEEB_CLUSTER", "1" if request.param == "cluster" else "0")
    return request.param


@pytest.mark.parametrize("bulk", [True, False])
@pytest.mark.parametrize("shape", SHAPES)
def test_parity_shapes(shape, bulk, kernel_path):
    _cabi.workspace().set_option("bulk", bulk)
    try:
        h, w = shape
        for maxval, bd in ((255, 8), (4095, 12), (65535, 16)):
            for gen, seed in ((synth_image, 3), (synth_saturated, 4)):
                img = gen(h, w, maxval, seed)
                for T in (1, 3, 20):
                    _check_one(img, T, bd, 0.9, seed + T)
    finally:
        _cabi.workspace().set_option("bulk", True)


def test_parity_reference_fixtures(golden_images, kernel_path):
    for name, bd in (("pe", 12), ("pe", 16), ("torax", 8)):
        for T in (1, 4, 16):
            _check_one(golden_images[name], T, bd, 0.97, 7)
    # SURVEY.md Appendix A reference point
    pe = golden_images["pe"]
    cap = _cap(pe, 1, 12)
    bits = np.random.default_rng(0).integers(0, 2, cap).astype(np.uint8)
    _, _, info = pee.pee_embed(pe, np.packbits(bits), 1, 12, n_bits=cap)
    assert info["n_flagged"] == 1127


def test_full_capacity_and_empty_payload(kernel_path):
    img = synth_image(96, 160, 4095, 9)
    for T in (1, 5):
        # a payload longer than any capacity, truncated to what fits: every carrier takes a real bit
        big = random_payload(img.size, 5)
        m0, lm0, i0 = PC.embed(img, big, img.size, T, 12)
        cap = i0["capacity"]
        rows = pee.pee_sweep(img, big, [T], 12, n_bits=img.size)
        assert rows[0]["capacity"] == cap and rows[0]["sse"] == i0["sse"] and rows[0]["n_flagged"] == i0["n_flagged"]
        m1, lm1, i1 = pee.pee_embed(img, big, T, 12, n_bits=cap)
        assert np.array_equal(m1, m0) and np.array_equal(lm1, lm0) and i1["capacity"] == cap
        p1, r1 = pee.pee_extract(m1, lm1, T, cap, 12)
        assert np.array_equal(r1, img) and np.array_equal(p1, np.packbits(np.unpackbits(big)[:cap]))
        with pytest.raises(ValueError):
            pee.pee_embed(img, big, T, 12, n_bits=cap + 1)
        with pytest.raises(ValueError):
            pee.pee_extract(m1, lm1, T, cap + 1, 12)
    m, lm, info = pee.pee_embed(img, b"", 3, 12, n_bits=0)
    m0, lm0, i0 = PC.embed(img, np.zeros(1, np.uint8), 0, 3, 12)
    i0.pop("status")
    assert np.array_equal(m, m0) and np.array_equal(lm, lm0) and info == i0
    p, r = pee.pee_extract(m, lm, 3, 0, 12)
    assert p.size == 0 and np.array_equal(r, img)


def test_tiny_images_and_errors():
    for shape in ((1, 1), (2, 9), (9, 2), (1, 50)):
        img = synth_image(shape[0], shape[1], 255, 1)
        m, lm, info = pee.pee_embed(img, b"", 2, 8, n_bits=0)
        assert info["capacity"] == 0 and np.array_equal(m, img) and not lm.any()
        with pytest.raises(ValueError):
            pee.pee_embed(img, b"\xff", 2, 8, n_bits=1)
        p, r = pee.pee_extract(m, lm, 2, 0, 8)
        assert np.array_equal(r, img)
    with pytest.raises(ValueError):
        pee.pee_embed(np.zeros((8, 8), np.float32), b"", 1)
    with pytest.raises(ValueError):
        pee.pee_embed(np.zeros((8, 8), np.uint8), b"", 0)
    with pytest.raises(ValueError):
        pee.pee_embed(np.zeros((8, 8), np.uint8), b"", 200)
    with pytest.raises(ValueError):
        pee.pee_embed(np.zeros((2, 8, 8), np.uint8), b"", 1)


def test_inputs_not_mutated():
    img = synth_image(80, 96, 4095, 2)
    keep = img.copy()
    pay = random_payload(500, 1)
    keep_pay = pay.copy()
    m, lm, _ = pee.pee_embed(img, pay, 4, 12, n_bits=500)
    mk, lk = m.copy(), lm.copy()
    pee.pee_extract(m, lm, 4, 500, 12)
    assert np.array_equal(img, keep) and np.array_equal(pay, keep_pay) and np.array_equal(m, mk) and np.array_equal(lm, lk)


def test_histogram_and_auto_threshold():
    for maxval, bd in ((255, 8), (4095, 12), (65535, 16)):
        for gen in (synth_image, synth_saturated):
            img = gen(90, 131, maxval, 7)
            assert np.array_equal(pee.pee_histogram(img, bd), PN.error_histogram(img, bd))
    img = synth_image(120, 200, 4095, 11)
    for n_bits in (100, 3000, 9000):
        pay = random_payload(n_bits, n_bits)
        m1, lm1, i1 = pee.pee_embed(img, pay, None, 12, n_bits=n_bits)
        m0, lm0, i0 = PN.pee_embed(img, pay, None, 12, n_bits=n_bits)
        assert i1 == i0 and np.array_equal(m1, m0) and np.array_equal(lm1, lm0)
    with pytest.raises(ValueError):
        pee.pee_embed(img, random_payload(img.size, 1), None, 12, n_bits=img.size)


def test_sweep_matches_oracle():
    img = synth_image(128, 192, 65535, 5)
    pay = random_payload(img.size, 3)
    Ts = list(range(1, 33))
    got = pee.pee_sweep(img, pay, Ts, 16, n_bits=img.size)
    want = PN.pee_sweep(img, pay, Ts, 16, n_bits=img.size)
    for g, w_ in zip(got, want):
        assert {k: g[k] for k in ("T", "capacity", "cap0", "cap1", "n_flagged", "sse")} == \
               {k: w_[k] for k in ("T", "capacity", "cap0", "cap1", "n_flagged", "sse")}
        assert g["mse"] == w_["mse"] and abs(g["psnr"] - w_["psnr"]) <= 1e-9 * abs(w_["psnr"])


def test_batch_matches_oracle_per_unit(kernel_path):
    imgs = synth_batch(7, 130, 264, 4095, 40)
    Ts = np.array([1, 2, 3, 4, 5, 6, 7], np.int32)
    nb = np.array([0, 50, 400, 1000, 2000, 33, 777], np.int64)
    stride = int(((nb + 7) // 8).max())
    pays = np.zeros((7, stride), np.uint8)
    for u in range(7):
        p = random_payload(int(nb[u]), u)
        pays[u, :p.size] = p
    marked, lm, info = pee.pee_embed_batch(imgs, pays, nb, Ts, 12)
    assert (info[:, 7] == 0).all()
    for u in range(7):
        m0, lm0, i0 = PC.embed(imgs[u], pays[u], int(nb[u]), int(Ts[u]), 12)
        assert np.array_equal(marked[u], m0) and np.array_equal(lm[u], lm0)
        assert [int(v) for v in info[u, :7]] == [i0[k] for k in ("T", "n_bits", "capacity", "cap0", "cap1", "n_flagged", "sse")]
    out, rec, xinfo = pee.pee_extract_batch(marked, lm, Ts, nb, 12)
    assert np.array_equal(rec, imgs) and np.array_equal(out, pays) and (xinfo[:, 7] == 0).all()
    assert np.array_equal(xinfo[:, 2], info[:, 2])


def test_batch_auto_threshold_matches_oracle_per_unit():
    """T=None for a batch: every unit's threshold is chosen on the device (histogram estimate, then only the units that
    fall short are embedded again at T + 1) -- same T, marked image, location map and statistics as the oracle's
    per-image search; a unit no threshold can hold keeps the capacity status without disturbing the others."""
    for maxval, bd, gen in ((4095, 12, synth_batch), (255, 8, synth_batch), (65535, 16, None)):
        n, h, w = 9, 150, 272
        if gen is None:
            imgs = np.stack([synth_saturated(h, w, maxval, 60 + u) for u in range(n)])
        else:
            imgs = gen(n, h, w, maxval, 50)
        caps = [PC.embed(imgs[u], np.zeros(h * w // 8 + 8, np.uint8), 0, 1 << (bd - 1), bd)[2]["capacity"] for u in range(n)]
        fr = [0.0, 0.02, 0.2, 0.45, 0.6, 0.75, 0.9, 0.97, 2.0]
        nb = np.array([int(caps[u] * fr[u]) for u in range(n)], np.int64)
        stride = int(((nb + 7) // 8).max()) + 4
        stride += -stride % 4
        pays = np.zeros((n, stride), np.uint8)
        for u in range(n):
            p = random_payload(int(nb[u]), 300 + u)
            pays[u, :p.size] = p
        marked, lm, info = pee.pee_embed_batch(imgs, pays, nb, None, bd)
        fits = 0
        for u in range(n):
            try:
                m0, lm0, i0 = PN.pee_embed(imgs[u], pays[u], None, bd, n_bits=int(nb[u]))
            except ValueError:
                assert int(info[u, 7]) == pee.PEEB_E_CAPACITY, u
                continue
            fits += 1
            assert int(info[u, 7]) == 0, (u, info[u])
            assert [int(v) for v in info[u, :7]] == [i0[k] for k in ("T", "n_bits", "capacity", "cap0", "cap1", "n_flagged", "sse")], u
            assert np.array_equal(marked[u], m0) and np.array_equal(lm[u], lm0), u
        assert fits >= 7
        ok = info[:, 7] == 0
        out, rec, xinfo = pee.pee_extract_batch(marked[ok], lm[ok], info[ok, 0].astype(np.int32), nb[ok], bd)
        assert np.array_equal(rec, imgs[ok]) and (xinfo[:, 7] == 0).all()
        for k, u in enumerate(np.flatnonzero(ok)):
            assert np.array_equal(np.unpackbits(out[k])[:nb[u]], np.unpackbits(pays[u])[:nb[u]]), u
    with pytest.raises(ValueError):
        pee.pee_embed_batch(imgs[0], pays, nb, None, bd, shared_cover=True)


def test_large_batch_of_small_images_matches_oracle_per_unit():
    """Many units with several bands each: the payload assembly takes several pieces per block, the units'
    summaries come from their last bands, the output rows are cleared by the kernel (dirty buffers passed in)."""
    n, h, w, bd = 320, 130, 72, 12
    imgs = synth_batch(n, h, w, 4095, 90)
    rng = np.random.default_rng(5)
    Ts = rng.integers(1, 9, n).astype(np.int32)
    caps = np.array([PC.embed(imgs[u], np.zeros(h * w // 8 + 8, np.uint8), 0, int(Ts[u]), bd)[2]["capacity"] for u in range(n)])
    # (the capacity of pass 1 depends on the payload: stay a little under the zero-payload capacity)
    nb = np.where(rng.integers(0, 4, n) == 0, (caps * 0.93).astype(np.int64), (caps * 0.9 * rng.random(n)).astype(np.int64)).astype(np.int64)
    nb[::17] = 0
    stride = int(((nb + 7) // 8).max()) + 3
    pays = np.zeros((n, stride), np.uint8)
    for u in range(n):
        p = random_payload(int(nb[u]), 1000 + u)
        pays[u, :p.size] = p
    marked, lm, info = pee.pee_embed_batch(imgs, pays, nb, Ts, bd)
    for u in range(n):
        m0, lm0, i0 = PC.embed(imgs[u], pays[u], int(nb[u]), int(Ts[u]), bd)
        assert int(info[u, 7]) == i0["status"] == 0, u
        assert np.array_equal(marked[u], m0) and np.array_equal(lm[u], lm0), u
        assert [int(v) for v in info[u, :7]] == [i0[k] for k in ("T", "n_bits", "capacity", "cap0", "cap1", "n_flagged", "sse")], u
    out = np.full((n, stride), 0xA5, np.uint8)
    rec = np.full_like(imgs, 7)
    out, rec, xinfo = pee.pee_extract_batch(marked, lm, Ts, nb, bd, out_payload=out, out_recovered=rec)
    assert np.array_equal(rec, imgs) and (xinfo[:, 7] == 0).all() and np.array_equal(xinfo[:, 2], info[:, 2])
    for u in range(n):
        k = int(nb[u])
        assert np.array_equal(np.unpackbits(out[u])[:k], np.unpackbits(pays[u])[:k]), u
        assert not np.unpackbits(out[u])[k:((k + 7) // 8) * 8].any(), u


@pytest.mark.parametrize("cfg", [
    dict(name="ct512", n=64, h=512, w=512, maxval=65535, bd=16, T=96),      # BASELINE configs[1]/[2] slice shape
    dict(name="dx3000", n=3, h=3000, w=3000, maxval=4095, bd=12, T=12),     # BASELINE configs[3] image shape
    dict(name="sweep2048", n=2, h=2048, w=2048, maxval=65535, bd=16, T=300), # BASELINE configs[4] image shape
])
def test_full_size_roundtrip_and_oracle(cfg):
    """BASELINE.json shapes: max-capacity embed -> extract is the identity on
    image and payload for the whole batch, and EVERY image is checked bit for bit
    (marked image, location map, statistics) against the C oracle."""
    n, h, w, bd, T = cfg["n"], cfg["h"], cfg["w"], cfg["bd"], cfg["T"]
    imgs = synth_batch(n, h, w, cfg["maxval"], 100)
    stride = (h * w + 7) // 8
    rng = np.random.default_rng(1)
    pays = rng.integers(0, 256, size=(n, stride), dtype=np.uint8)
    big = np.full(n, h * w, np.int64)
    _, _, info = pee.pee_embed_batch(imgs, pays, big, T, bd, want_marked=False, want_lm=False)
    cap = info[:, 2].copy()
    assert (info[:, 7] == _cabi.PEEB_E_CAPACITY).all() and (cap > 0).all()
    marked, lm, info2 = pee.pee_embed_batch(imgs, pays, cap, T, bd)
    assert (info2[:, 7] == 0).all() and np.array_equal(info2[:, 2], cap)
    out, rec, xinfo = pee.pee_extract_batch(marked, lm, T, cap, bd)
    assert np.array_equal(rec, imgs)
    assert np.array_equal(xinfo[:, 2], cap) and (xinfo[:, 7] == 0).all()
    for u in range(n):
        nbytes, rem = int(cap[u]) // 8, int(cap[u]) % 8
        assert np.array_equal(out[u, :nbytes], pays[u, :nbytes])
        if rem:
            assert int(out[u, nbytes]) == (int(pays[u, nbytes]) & ((0xFF00 >> rem) & 0xFF))
    m0, lm0, i0 = PC.embed_batch(imgs, np.pad(pays, ((0, 0), (0, 8))), cap, T, bd)
    assert np.array_equal(marked, m0), f"marked images differ in units {np.flatnonzero((marked != m0).reshape(n, -1).any(axis=1))[:8]}"
    assert np.array_equal(lm, lm0), "location maps differ"
    assert np.array_equal(info2[:, :7], i0[:, :7]), "statistics differ"


def test_sweep_pairs_series_matches_oracle():
    """BASELINE configs[4] in miniature: every (image, T) pair of a small series."""
    from codec_tcc_b200 import shard
    imgs = synth_batch(3, 96, 160, 65535, 70)
    pays = np.random.default_rng(2).integers(0, 256, (3, 96 * 160 // 8), dtype=np.uint8)
    Ts = [1, 8, 64, 300]
    table = shard.sweep_sharded(imgs, pays, Ts, 16)      # world size 1 here: the whole grid
    assert table.shape == (12, 10) and np.array_equal(table[:, 0], np.repeat(np.arange(3), 4))
    for row in table:
        u, T = int(row[0]), int(row[1])
        _, _, i0 = PC.embed(imgs[u], pays[u], 96 * 160, T, 16)
        assert [int(v) for v in row[3:8]] == [i0[k] for k in ("capacity", "cap0", "cap1", "n_flagged", "sse")]
