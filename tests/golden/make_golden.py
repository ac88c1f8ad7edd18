"""Generate tests/golden/{fixtures.npz, reference_golden.json} by RUNNING THE
UNMODIFIED REFERENCE (``/root/reference/src/{codec,mse}.py`` through
``oracle/ref_import.py``) on its own two images and on seeded synthetic
inputs.  Run in the build container only (the reference tree is not on the GPU
box); the outputs are committed.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from codec_tcc_b200.synth import synth_image, synth_saturated  # noqa: E402
from oracle import ref_import  # noqa: E402
import _golden_checks as GC  # noqa: E402  (input builders shared with the tests)

HERE = os.path.dirname(os.path.abspath(__file__))
MAIN_MESSAGE = "Mensagem de teste para esteganografia!"  # src/codec.py:863


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rand_bitstring(n, seed):
    rng = np.random.default_rng(seed)
    return "".join("1" if b else "0" for b in rng.integers(0, 2, n).tolist())


def lsb_case(codec, mse, img, bits, beta, sbs, align, embedder):
    with ref_import.quiet():
        g, l = codec.adaptive_modalities_decomposition(img, beta=beta)
    s = len(l)
    if embedder == "hybrid":
        sp, bm, used, lens, idx = codec.lsb_embed_block_then_multiplane(
            l, bits, search_block_size=sbs, align_across_planes=align)
    else:
        sp, bm, used, lens, idx = codec.lsb_embed_multi_plane(l, bits)
    stego = codec.merge_modalities(g, sp)
    meta = {"s": s, "segments_indices": idx, "segments_lengths": lens}
    decoded = codec.decode_message(codec.extract_local_planes(stego, s), [b.ravel() for b in bm], meta)
    an = mse.AnalisadorMSE()
    with ref_import.quiet():
        m, rng_ = an.calcular_mse(img, stego)
        ssim = an.calcular_ssim_simples(img, stego)
    psnr = an.calcular_psnr(m, rng_)
    a = np.array(img, dtype=np.float64)
    b = np.array(stego, dtype=np.float64)
    return {
        "beta": beta, "sbs": sbs, "align": align, "embedder": embedder, "n_bits": len(bits),
        "s": s, "total_used": int(used), "segments_lengths": [int(v) for v in lens],
        "segment_indices": [int(v) for v in idx],
        "stego_dtype": str(stego.dtype), "stego_sha": sha(stego), "bitmaps_sha": sha(np.stack(bm)),
        "planes_sha": sha(np.stack(sp)), "decoded_len": len(decoded),
        "decoded_sha": hashlib.sha256(decoded.encode("utf-8")).hexdigest(),
        "decoded_head_hex": decoded.encode("utf-8")[:48].hex(),
        "px_changed": int(np.sum(a != b)), "mse": float(m), "max_range": float(rng_),
        "psnr": float(psnr), "ssim": float(ssim),
        "mean_abs": float(np.mean(np.abs(a - b))), "max_abs": float(np.max(np.abs(a - b))),
    }


def main():
    codec, mse = ref_import.codec(), ref_import.mse()
    pe = ref_import.read_fixture_pixels("pe")
    tx = ref_import.read_fixture_pixels("torax")
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), pe=pe, torax=tx)

    images = {
        "pe": pe, "torax": tx,
        "synth16_257x301": synth_image(257, 301, 65535, 11),
        "synth12_300x200": synth_image(300, 200, 4095, 12),
        "synth8_129x70": synth_image(129, 70, 255, 13),
        "sat12_96x160": synth_saturated(96, 160, 4095, 14),
    }
    out = {"_generator": "tests/golden/make_golden.py (unmodified reference via oracle/ref_import.py)",
           "numpy": np.__version__, "images": {}, "segments": {}, "scalars": {}}

    for name, img in images.items():
        rec = {"shape": list(img.shape), "dtype": str(img.dtype), "sha": sha(img)}
        rec["entropy"] = float(codec.calculate_entropy(img))
        nbits = img.dtype.itemsize * 8
        mi = []
        for i in range(nbits):
            mi.append(float(codec.calculate_mutual_information((img >> i) & 1, img)))
        rec["mi"] = mi
        rec["split"] = {}
        for beta in (0.2, 0.4, 0.6, 0.8, 0.95):
            with ref_import.quiet():
                g, l = codec.adaptive_modalities_decomposition(img, beta=beta)
            rec["split"][str(beta)] = len(l)
        cases = [lsb_case(codec, mse, img, codec.message_to_bits(MAIN_MESSAGE), 0.4, 16, False, "hybrid")]
        npx = img.size
        big = rand_bitstring(min(60000, npx // 2), 5)
        cases.append(lsb_case(codec, mse, img, big, 0.6, 8, False, "hybrid"))
        cases.append(lsb_case(codec, mse, img, big, 0.6, 8, True, "hybrid"))
        cases.append(lsb_case(codec, mse, img, big, 0.8, 16, False, "multi"))
        # payload so large that the biggest segments exceed the image (clamped, src/codec.py:464) and wrap
        wrap = rand_bitstring(int(npx * 4), 6)
        cases.append(lsb_case(codec, mse, img, wrap, 0.8, 16, False, "hybrid"))
        cases.append(lsb_case(codec, mse, img, wrap, 0.8, 16, False, "multi"))
        rec["lsb_cases"] = cases
        out["images"][name] = rec

    for s in (1, 2, 3, 5, 8, 12, 16):
        for total in (0, 1, 3, 7, 304, 1000, 65537):
            planes = [None] * s
            segs, sizes, order = codec.distribute_message_segments(planes, "0" * total)
            out["segments"][f"{s}:{total}"] = {"sizes": sizes, "order": order,
                                               "seg_lens": [len(x) for x in segs]}

    an = mse.AnalisadorMSE()
    sc = out["scalars"]
    with ref_import.quiet():
        sc["mse_norm_small"] = [float(v) for v in an.calcular_mse([[10, 20], [30, 40]], [[10, 20], [30, 41]])]
        a = synth_image(120, 90, 4095, 21)
        b = a.copy(); b[5, 7] += 900; b[60:70, 10:50] ^= 3
        sc["mse_norm_synth"] = [float(v) for v in an.calcular_mse(a, b)]
        sc["ssim_norm_synth"] = float(an.calcular_ssim_simples(a, b))
        sc["mse_same"] = [float(v) for v in an.calcular_mse(a, a)]
        sc["ssim_same"] = float(an.calcular_ssim_simples(a, a))
    sc["psnr_zero"] = an.calcular_psnr(0.0, 4095)
    sc["psnr_default"] = float(an.calcular_psnr(1.0))
    sc["psnr_4095"] = float(an.calcular_psnr(0.37, 4095))
    sc["message_bits_hex"] = codec.message_to_bits("Olá, DICOM ✓")
    sc["header_hex"] = None
    with ref_import.quiet():
        sc["header_hex"] = codec.create_header("jxl", 5, [1966, 1256, 706, 314, 78], [3, 1, 2, 4, 0],
                                               1234, 64, 64, 0, False).hex()

    # round 2: the general branch of calculate_mutual_information and the float64 route of the metrics
    sc["mi_general"] = {name: float(codec.calculate_mutual_information(plane, img)) for name, plane, img in GC.general_mi_cases()}
    sc["float_metrics"] = {}
    for name, x, y in GC.float_metric_cases():
        with ref_import.quiet():
            m, r = an.calcular_mse(x, y)
            ss = an.calcular_ssim_simples(x, y)
        fx, fy = np.array(x, dtype=np.float64), np.array(y, dtype=np.float64)
        sc["float_metrics"][name] = {"mse": float(m), "max_range": float(r), "ssim": float(ss),
                                     "mean_abs": float(np.mean(np.abs(fx - fy))), "max_abs": float(np.max(np.abs(fx - fy))),
                                     "changed": int(np.sum(fx != fy))}

    def enc(o):
        if isinstance(o, float) and o == float("inf"):
            return "inf"
        raise TypeError

    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(out, f, indent=1, default=enc, allow_nan=True)
    print("wrote", len(json.dumps(out, allow_nan=True)), "bytes of golden data")


if __name__ == "__main__":
    main()
