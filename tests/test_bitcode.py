"""N2: the "PBR1" coding of the side bitmaps -- bit packing + zero-run elimination on the GPU,
in place of the reference's zlib over one byte per pixel (src/codec.py:888-889, :820-821).

CPU part: the numpy restatement (oracle/bitcode_numpy.py) round-trips, its format is frozen by a known
answer, and the container functions keep reading the reference's zlib blobs.  GPU part (through the C ABI):
the encoder's bytes equal the restatement's, the decoder restores the reference's own bitmaps of the
golden LSB cases, corrupt blobs are refused."""
import hashlib
import zlib

import numpy as np
import pytest

from codec_tcc_b200 import container
from oracle import bitcode_numpy as BN

SIZES = (0, 1, 7, 8, 31, 32, 33, 1023, 1024, 1025, 32767, 32768, 32769, 100_003, 1_048_576 + 5)


def _map(n, density, seed):
    rng = np.random.default_rng(seed)
    return ((rng.random(n) < density) * rng.integers(1, 256, n)).astype(np.uint8)


def test_oracle_round_trip_and_forms():
    for n in SIZES:
        for density in (0.0, 0.0005, 0.3, 1.0):
            a = _map(n, density, n + 1)
            blob = BN.encode(a)
            assert np.array_equal(BN.decode(blob, n), (a != 0).astype(np.uint8))
            packed = np.packbits(a != 0)
            assert BN.encode(packed, packed=True, n=n) == blob
            assert np.array_equal(BN.decode(blob, n, packed=True), packed)


def test_oracle_format_known_answers():
    # an empty map of 36 M elements is its header + the top level only
    assert len(BN.encode(np.zeros(36_000_000, np.uint8))) == 24 + 4 * 2 * 1099
    # one set element: header, L2 (1 word), C0 (1 word: one non-zero L0 word), one L1 word, one L0 word;
    # element 9 = bit 6 of packed byte 1
    blob = BN.encode(np.eye(1, 40, 9, dtype=np.uint8))
    assert blob == b"PBR1" + bytes(4) + (40).to_bytes(8, "little") + (1).to_bytes(4, "little") * 2 \
        + (1).to_bytes(4, "little") + (1).to_bytes(4, "little") + (1).to_bytes(4, "little") + (0x4000).to_bytes(4, "little")
    a = _map(70_001, 0.01, 5)
    frozen = BN.encode(a)   # frozen with the first version of the format: a change of layout must be deliberate
    assert len(frozen) == 2864 and hashlib.sha256(frozen).hexdigest()[:16] == "b4e7fc60f53caa49"
    with pytest.raises(ValueError):
        BN.decode(blob[:-4], 40)
    with pytest.raises(ValueError):
        BN.decode(blob, 41)
    with pytest.raises(ValueError):
        BN.decode(b"XXXX" + blob[4:], 40)


def test_container_still_reads_reference_blobs():
    maps = [_map(500, 0.2, k).reshape(20, 25) for k in range(3)]
    blob = container.pack_bitmaps(maps)
    assert blob == zlib.compress(np.stack(maps).tobytes())
    back = container.unpack_bitmaps(blob, 3)
    assert all(np.array_equal(b, m.ravel()) for b, m in zip(back, maps))
    with pytest.raises(ValueError):
        container.pack_bitmaps(maps, coding="lz4")


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_encoder_equals_restatement_and_decoder_inverts():
    for n in SIZES:
        for density in (0.0, 0.0005, 0.3, 1.0):
            a = _map(n, density, 7 * n + 3)
            want = BN.encode(a)
            got = container.encode_bitmap(a)
            assert got == want, (n, density)
            assert np.array_equal(container.decode_bitmap(got, n), (a != 0).astype(np.uint8))
            packed = np.packbits(a != 0)
            assert container.encode_bitmap(packed, packed=True, n=n) == want
            assert np.array_equal(container.decode_bitmap(got, n, packed=True), packed)


@pytest.mark.gpu
def test_gpu_unaligned_inputs_and_outputs():
    base = _map(200_000, 0.05, 9)
    for off in (1, 3, 8):
        a = base[off:off + 150_001]
        assert container.encode_bitmap(a) == BN.encode(a)
    import torch
    from codec_tcc_b200 import device as D
    t = torch.from_numpy(base).cuda()
    for off in (0, 1, 5, 16):
        v = t[off:off + 131_073]
        blob = D.bitmap_encode_device(v)
        assert bytes(blob.cpu().numpy()) == BN.encode(base[off:off + 131_073])
        out = torch.empty(131_073 + 16, dtype=torch.uint8, device="cuda")
        for ooff in (0, 1):
            D.bitmap_decode_device(blob, 131_073, out=out[ooff:ooff + 131_073])
            assert np.array_equal(out[ooff:ooff + 131_073].cpu().numpy(), (base[off:off + 131_073] != 0).astype(np.uint8))


@pytest.mark.gpu
def test_gpu_bitmaps_of_the_reference_golden_cases(golden, golden_images):
    """The reference's own XOR bitmaps (sha256 pinned to the unmodified reference) survive the device coding, and
    the container's blob step yields exactly what the reference's zlib step yields after decoding."""
    from codec_tcc_b200 import codec
    from _golden_checks import case_bits

    seen = 0
    for name, entry in golden["images"].items():
        img = golden_images[name]
        for case in entry["lsb_cases"][:3]:
            bits = case_bits(case, img, codec)
            g, l = codec.adaptive_modalities_decomposition(img, beta=case["beta"])
            if case["embedder"] == "hybrid":
                sp, bm, *_ = codec.lsb_embed_block_then_multiplane(l, bits, search_block_size=case["sbs"],
                                                                   align_across_planes=case["align"])
            else:
                sp, bm, *_ = codec.lsb_embed_multi_plane(l, bits)
            assert hashlib.sha256(np.stack(bm).tobytes()).hexdigest() == case["bitmaps_sha"]
            s = len(bm)
            blob = container.pack_bitmaps(bm, coding="pbr")
            assert blob == BN.encode(np.stack(bm))
            ref = container.unpack_bitmaps(container.pack_bitmaps(bm), s)   # the reference's zlib path
            got = container.unpack_bitmaps(blob, s)
            assert len(got) == s and all(np.array_equal(a, b) for a, b in zip(got, ref))
            assert len(blob) <= np.stack(bm).size // 8 * 1.04 + 64
            seen += 1
    assert seen >= 6


@pytest.mark.gpu
def test_gpu_location_map_of_a_pee_embed():
    from codec_tcc_b200 import pee
    from codec_tcc_b200.synth import random_payload, synth_saturated

    img = synth_saturated(257, 301, 4095, 4)
    marked, lm, info = pee.pee_embed(img, random_payload(3000, 1), 3, 12, n_bits=3000)
    assert info["n_flagged"] > 0
    blob = container.encode_bitmap(lm, packed=True)
    assert blob == BN.encode(lm, packed=True)
    assert np.array_equal(container.decode_bitmap(blob, lm.size * 8, packed=True).reshape(lm.shape), lm)
    assert len(blob) < lm.size


@pytest.mark.gpu
def test_gpu_decoder_refuses_corrupt_blobs():
    a = _map(100_000, 0.02, 2)
    blob = bytearray(container.encode_bitmap(a))
    with pytest.raises(ValueError):
        container.decode_bitmap(bytes(blob), 100_001)
    with pytest.raises(ValueError):
        container.decode_bitmap(bytes(blob[:-4]), 100_000)
    with pytest.raises(ValueError):
        container.decode_bitmap(b"ZLIB" + bytes(blob[4:]), 100_000)
    bad = bytearray(blob)
    bad[24] ^= 0xFF          # the top level no longer matches the counts in the header
    with pytest.raises(ValueError):
        container.decode_bitmap(bytes(bad), 100_000)
    bad = bytearray(blob)
    bad[24 + 4 * 4] ^= 0x01  # a block count (C0) that disagrees with the level-1 words
    with pytest.raises(ValueError):
        container.decode_bitmap(bytes(bad), 100_000)
    assert np.array_equal(container.decode_bitmap(bytes(blob), 100_000), (a != 0).astype(np.uint8))


def test_oracle_properties_random_maps():
    """Property test of the restatement (hypothesis): decode(encode(m)) == m for arbitrary lengths and densities, the packed
    and the byte form give the same blob, the blob never exceeds its bound and shrinks with sparsity."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 70_000), st.floats(0.0, 1.0), st.integers(0, 2 ** 31 - 1))
    def prop(n, density, seed):
        a = _map(n, density ** 3, seed)
        blob = BN.encode(a)
        n0 = (n + 31) // 32; n1 = (n0 + 31) // 32; n2 = (n1 + 31) // 32
        assert len(blob) <= 24 + 4 * (2 * n2 + n1 + n0)
        assert np.array_equal(BN.decode(blob, n), (a != 0).astype(np.uint8))
        assert BN.encode(np.packbits(a != 0), packed=True, n=n) == blob
        if not a.any():
            assert len(blob) == 24 + 8 * n2

    prop()
