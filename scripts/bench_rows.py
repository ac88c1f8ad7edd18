"""Measures the non-headline rows of SURVEY.md section 8 (a1-a9) on one B200:
each C-ABI kernel device-resident (CUDA events, achieved algorithmic GB/s vs the
measured copy peak) and each numpy-API call end to end, next to the CPU oracle
(the numpy restatement of the reference) timed on the same inputs.

    python scripts/bench_rows.py > gpurun_out/rows.json
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from codec_tcc_b200 import _cabi, codec, mse  # noqa: E402
from codec_tcc_b200.synth import synth_image  # noqa: E402
from oracle import codec_numpy as OC  # noqa: E402
from oracle import mse_numpy as OM  # noqa: E402

L = _cabi.lib()
ws = _cabi.workspace(0)
dev = torch.device("cuda", 0)
peak = 6557.1
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    peak = float(json.load(open(pp))["hbm_gbs"])


def dev_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the calls queue up behind ~10 ms of spinning: kernels of a few microseconds are then timed at the device's
    # pace, not at the rate the host can enqueue them (a call costs the host 30-40 us through ctypes)
    torch.cuda._sleep(20_000_000)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def wall(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def stream():
    return torch.cuda.current_stream(dev).cuda_stream


rows = []
h = w = 3000
n_img = 16
imgs = np.stack([synth_image(h, w, 4095, 50 + k) for k in range(n_img)])
stego = imgs.copy()
stego[:, ::3, ::5] ^= 1
d_a = torch.from_numpy(imgs.view(np.int16)).to(dev)
d_b = torch.from_numpy(stego.view(np.int16)).to(dev)
npx = h * w

# ---- a1-a4: moments (16 images of 3000x3000: 288 MB per operand, larger than L2)
d_out = torch.empty((n_img, 12), dtype=torch.int64, device=dev)
t = dev_time(lambda: _cabi.check(L.peeb_moments_batch(ws.handle, d_a.data_ptr(), d_b.data_ptr(), npx, 2, n_img, npx, npx,
                                                      d_out.data_ptr(), stream())))
rows.append({"row": "a1-a4", "kernel": "moments_kernel", "workload": f"{n_img} x {h}x{w} u16 pairs",
             "device_ms": t * 1e3, "mpixel_s": n_img * npx / t / 1e6, "algorithmic_gb_s": 4 * n_img * npx / t / 1e9,
             "frac_of_measured_peak": 4 * n_img * npx / t / 1e9 / peak})
t = dev_time(lambda: _cabi.check(L.peeb_sse_batch(ws.handle, d_a.data_ptr(), d_b.data_ptr(), npx, 2, n_img, npx, npx,
                                                  d_out.data_ptr(), stream())))
rows.append({"row": "a1", "kernel": "moments_kernel<FULL=false> (SSE + maxima: calcular_mse)", "workload": f"{n_img} x {h}x{w} u16 pairs",
             "device_ms": t * 1e3, "mpixel_s": n_img * npx / t / 1e6, "algorithmic_gb_s": 4 * n_img * npx / t / 1e9,
             "frac_of_measured_peak": 4 * n_img * npx / t / 1e9 / peak})
an = mse.AnalisadorMSE()
t_api = wall(lambda: an.calcular_mse(imgs[0], stego[0]))
t_cpu = wall(lambda: OM.calcular_mse(imgs[0], stego[0]), reps=1)
rows[-1].update(api_ms=t_api * 1e3, api_mpixel_s=npx / t_api / 1e6, cpu_oracle_ms=t_cpu * 1e3,
                cpu_oracle_mpixel_s=npx / t_cpu / 1e6, api="AnalisadorMSE.calcular_mse (one image pair, pageable numpy arrays)")
t_api = wall(lambda: an.analisar_par_arrays(imgs[0], stego[0]))
t_cpu = wall(lambda: (OM.calcular_mse(imgs[0], stego[0]), OM.calcular_ssim_simples(imgs[0], stego[0]),
                      OM.difference_stats(imgs[0], stego[0])), reps=1)
rows[-2].update(api_ms=t_api * 1e3, api_mpixel_s=npx / t_api / 1e6, cpu_oracle_ms=t_cpu * 1e3,
                cpu_oracle_mpixel_s=npx / t_cpu / 1e6, api="AnalisadorMSE.analisar_par_arrays (mse+psnr+ssim+stats, one image pair)")

# ---- a5: histogram + plane counts, plane split
d_hist = torch.empty(65536, dtype=torch.int32, device=dev)
d_ones = torch.empty(16, dtype=torch.int64, device=dev)
big = d_a.reshape(-1)
t = dev_time(lambda: _cabi.check(L.peeb_hist_planes(ws.handle, big.data_ptr(), big.numel(), 2, d_hist.data_ptr(),
                                                    d_ones.data_ptr(), stream())))
rows.append({"row": "a5", "kernel": "hist_kernel+hist_to_planes", "workload": f"{n_img * npx / 1e6:.0f} Mpixel u16 (12-bit values)",
             "device_ms": t * 1e3, "mpixel_s": big.numel() / t / 1e6, "algorithmic_gb_s": 2 * big.numel() / t / 1e9,
             "frac_of_measured_peak": 2 * big.numel() / t / 1e9 / peak})
t_api = wall(lambda: codec.adaptive_modalities_decomposition(imgs[0], beta=0.8))
t_cpu = wall(lambda: OC.adaptive_modalities_decomposition(imgs[0], beta=0.8), reps=1)
rows[-1].update(api_ms=t_api * 1e3, api_mpixel_s=npx / t_api / 1e6, cpu_oracle_ms=t_cpu * 1e3,
                cpu_oracle_mpixel_s=npx / t_cpu / 1e6, api="adaptive_modalities_decomposition (one 3000x3000 image, 16 planes returned)")

d_planes = torch.empty((16, npx), dtype=torch.int16, device=dev)
t = dev_time(lambda: _cabi.check(L.peeb_planes_unpack(ws.handle, d_a.data_ptr(), npx, 2, 0, 16, d_planes.data_ptr(), stream())))
rows.append({"row": "a5/a8", "kernel": "planes_unpack_kernel", "workload": "3000x3000 u16 -> 16 planes", "device_ms": t * 1e3,
             "mpixel_s": npx / t / 1e6, "algorithmic_gb_s": (2 + 32) * npx / t / 1e9, "frac_of_measured_peak": (2 + 32) * npx / t / 1e9 / peak})
d_packed = torch.empty(npx, dtype=torch.int16, device=dev)
t = dev_time(lambda: _cabi.check(L.peeb_planes_pack(ws.handle, d_planes.data_ptr(), npx, 2, 16, d_packed.data_ptr(), stream())))
rows.append({"row": "a8", "kernel": "planes_pack_kernel", "workload": "16 planes -> 3000x3000 u16", "device_ms": t * 1e3,
             "mpixel_s": npx / t / 1e6, "algorithmic_gb_s": (2 + 32) * npx / t / 1e9, "frac_of_measured_peak": (2 + 32) * npx / t / 1e9 / peak})
g, l = codec.adaptive_modalities_decomposition(imgs[0], beta=0.8)
t_api = wall(lambda: codec.merge_modalities(g, l))
t_cpu = wall(lambda: OC.merge_modalities(g, l), reps=1)
rows[-1].update(api_ms=t_api * 1e3, api_mpixel_s=npx / t_api / 1e6, cpu_oracle_ms=t_cpu * 1e3, cpu_oracle_mpixel_s=npx / t_cpu / 1e6,
                api="merge_modalities (16 planes)")

# ---- a6/a7: tile moments + LSB embed (s planes in, s planes + s uint8 bitmaps out)
s = len(l)
d_sums = torch.empty((((h + 15) // 16) * ((w + 15) // 16), 2), dtype=torch.int64, device=dev)
t = dev_time(lambda: _cabi.check(L.peeb_tile_moments(ws.handle, d_planes.data_ptr(), h, w, 2, 16, d_sums.data_ptr(), stream())))
rows.append({"row": "a6", "kernel": "tile_moments_kernel", "workload": "3000x3000 plane, 16x16 tiles", "device_ms": t * 1e3,
             "mpixel_s": npx / t / 1e6, "algorithmic_gb_s": 2 * npx / t / 1e9, "frac_of_measured_peak": 2 * npx / t / 1e9 / peak})
pay_bits = int(npx * 2.0)
rng = np.random.default_rng(1)
bits = "".join("1" if b else "0" for b in rng.integers(0, 2, pay_bits).tolist())
d_out_planes = torch.empty((s, npx), dtype=torch.int16, device=dev)
d_bm = torch.empty((s, npx), dtype=torch.uint8, device=dev)
d_pay = torch.from_numpy(np.packbits(rng.integers(0, 2, pay_bits, dtype=np.uint8))).to(dev)
segs, sizes, order = codec.distribute_message_segments(l, bits)
start = np.zeros(s, np.int64); ln = np.zeros(s, np.int64); off = np.zeros(s, np.int64)
at = 0
for seg, p in zip(segs, order):
    ln[p] = min(len(seg), npx); off[p] = at; at += ln[p]
t = dev_time(lambda: _cabi.check(L.peeb_lsb_embed(ws.handle, d_planes.data_ptr(), npx, 2, s, start.ctypes.data, ln.ctypes.data,
                                                 off.ctypes.data, d_pay.data_ptr(), pay_bits, d_out_planes.data_ptr(),
                                                 d_bm.data_ptr(), stream())))
rows.append({"row": "a6/a7", "kernel": "lsb_embed_kernel", "workload": f"s={s} planes of 3000x3000 u16, {pay_bits / 1e6:.1f} Mbit payload",
             "device_ms": t * 1e3, "mpixel_s": npx / t / 1e6, "algorithmic_gb_s": s * 5 * npx / t / 1e9,
             "frac_of_measured_peak": s * 5 * npx / t / 1e9 / peak})
t_api = wall(lambda: codec.lsb_embed_block_then_multiplane(l, bits, search_block_size=16), reps=1)
t0 = time.perf_counter(); ref = OC.lsb_embed_block_then_multiplane(l, bits, search_block_size=16); t_cpu = time.perf_counter() - t0
got = codec.lsb_embed_block_then_multiplane(l, bits, search_block_size=16)
assert all(np.array_equal(a, b) for a, b in zip(ref[0] + ref[1], got[0] + got[1])) and ref[2:] == got[2:]
rows[-1].update(api_ms=t_api * 1e3, api_mpixel_s=npx / t_api / 1e6, cpu_oracle_ms=t_cpu * 1e3, cpu_oracle_mpixel_s=npx / t_cpu / 1e6,
                api="lsb_embed_block_then_multiplane (numpy planes in/out, '0'/'1' string payload); outputs equal the restatement")

# ---- a5-a8 chained: the reference's encode flow (decompose -> hybrid embed -> merge), device resident
t_api = wall(lambda: codec.embed_pipeline(imgs[0], bits, beta=0.8, search_block_size=16), reps=2)
t0 = time.perf_counter()
g0, l0 = OC.adaptive_modalities_decomposition(imgs[0], beta=0.8)
r_e = OC.lsb_embed_block_then_multiplane(l0, bits, search_block_size=16)
r_m = OC.merge_modalities(g0, r_e[0])
t_cpu = time.perf_counter() - t0
p_st, p_bm, p_meta = codec.embed_pipeline(imgs[0], bits, beta=0.8, search_block_size=16)
assert np.array_equal(p_st, r_m) and all(np.array_equal(a, b) for a, b in zip(p_bm, r_e[1])) and p_meta["segments_indices"] == r_e[4]
rows.append({"row": "a5-a8", "kernel": "hist + planes_unpack + tile_moments + lsb_embed + planes_pack (one upload, one download)",
             "workload": f"3000x3000 u16 (12-bit), {pay_bits / 1e6:.1f} Mbit payload, s={s}", "api_ms": t_api * 1e3,
             "api_mpixel_s": npx / t_api / 1e6, "cpu_oracle_ms": t_cpu * 1e3, "cpu_oracle_mpixel_s": npx / t_cpu / 1e6,
             "api": "codec.embed_pipeline (numpy image + '0'/'1' string in, stego image + uint8 bitmaps out) against the chained "
                    "restatements of adaptive_modalities_decomposition, lsb_embed_block_then_multiplane, merge_modalities; outputs equal"})

# ---- N2: the same flow with the bitmaps coded on the device (PBR1 blob instead of s*h*w bytes over PCIe), beside the
# reference's zlib blob step on the host (src/codec.py:887-889)
import zlib  # noqa: E402
from codec_tcc_b200 import container  # noqa: E402
from oracle import bitcode_numpy as BN  # noqa: E402
t_api = wall(lambda: codec.embed_pipeline(imgs[0], bits, beta=0.8, search_block_size=16, bitmaps_as="pbr"), reps=2)
_, blob, _ = codec.embed_pipeline(imgs[0], bits, beta=0.8, search_block_size=16, bitmaps_as="pbr")
assert blob == BN.encode(p_bm) and all(np.array_equal(a, b.ravel()) for a, b in zip(container.unpack_bitmaps(blob, len(p_bm)), p_bm))
t0 = time.perf_counter(); zb = zlib.compress(np.stack(p_bm).tobytes()); t_z = time.perf_counter() - t0
rows.append({"row": "a5-a8 + N2", "kernel": "... + pbr_pack / pbr_scan / pbr_compact (bitmaps coded in device memory)",
             "workload": f"3000x3000 u16 (12-bit), {pay_bits / 1e6:.1f} Mbit payload, s={s}", "api_ms": t_api * 1e3,
             "api_mpixel_s": npx / t_api / 1e6, "blob_bytes": len(blob), "zlib_blob_bytes": len(zb), "cpu_zlib_ms": t_z * 1e3,
             "api": "codec.embed_pipeline(bitmaps_as='pbr'): stego image + PBR1 blob out; the blob decodes to the plain call's bitmaps; "
                    "cpu_zlib_ms = the reference's zlib.compress(np.stack(bitmaps).tobytes()) on these bitmaps, 1 core"})
d_maps = torch.from_numpy(np.stack(p_bm).reshape(-1)).to(dev)
d_blob = torch.empty(int(L.peeb_bitmap_blob_bound(d_maps.numel())), dtype=torch.uint8, device=dev)
from codec_tcc_b200 import device as D  # noqa: E402
ws.prof_enable(True)
for _ in range(5):
    bl = D.bitmap_encode_device(d_maps, blob=d_blob)
    D.bitmap_decode_device(bl, d_maps.numel())
torch.cuda.synchronize()
prof = ws.prof_report(); ws.prof_enable(False)
for key, label in (("bitmap_encode", "pbr_pack + pbr_scan + pbr_compact"), ("bitmap_decode", "pbr_plan + pbr_expand")):
    ms = prof[key][0] / prof[key][1]
    rows.append({"row": "N2", "kernel": label, "workload": f"s={s} bitmaps of 3000x3000 ({d_maps.numel() / 1e6:.0f} MB, one byte per element)",
                 "device_ms": ms, "algorithmic_gb_s": d_maps.numel() * 1.125 / 1e9 / (ms * 1e-3),
                 "frac_of_measured_peak": d_maps.numel() * 1.125 / 1e9 / (ms * 1e-3) / peak,
                 "note": "algorithmic bytes: one byte per element read (written for decode) + at most one bit per element on the other side"})

# ---- a9: decode_message
meta = {"s": s, "segments_indices": got[4], "segments_lengths": got[3]}
t_api = wall(lambda: codec.decode_message(got[0], got[1], meta), reps=1)
t0 = time.perf_counter(); r0 = OC.decode_message(ref[0], ref[1], meta); t_cpu = time.perf_counter() - t0
assert r0 == codec.decode_message(got[0], got[1], meta)
rows.append({"row": "a9", "kernel": "compact_{count,scan,write}", "workload": f"s={s} planes 3000x3000", "api_ms": t_api * 1e3,
             "api_mpixel_s": npx / t_api / 1e6, "cpu_oracle_ms": t_cpu * 1e3, "cpu_oracle_mpixel_s": npx / t_cpu / 1e6,
             "api": "decode_message (outputs equal the restatement)"})

# ---- N4: the true inverse of the LSB path (recover the cover, read the message back)
stego_img = codec.merge_modalities(g, got[0])
d_stego = torch.from_numpy(stego_img.view(np.int16)).to(dev).reshape(-1)
d_bm2 = torch.from_numpy(np.stack([b.reshape(-1) for b in got[1]])).to(dev)
d_cov = torch.empty_like(d_stego)
t = dev_time(lambda: _cabi.check(L.peeb_lsb_recover(ws.handle, d_stego.data_ptr(), d_bm2.data_ptr(), npx, 2, s, d_cov.data_ptr(), stream())))
assert np.array_equal(d_cov.cpu().numpy().view(np.uint16).reshape(h, w), imgs[0])
rows.append({"row": "N4", "kernel": "lsb_recover_kernel", "workload": f"3000x3000 u16 stego + s={s} uint8 bitmaps -> cover",
             "device_ms": t * 1e3, "mpixel_s": npx / t / 1e6, "algorithmic_gb_s": (4 + s) * npx / t / 1e9,
             "frac_of_measured_peak": (4 + s) * npx / t / 1e9 / peak})
meta4 = {"s": s, "segments_indices": got[4], "segments_lengths": got[3], "hybrid": True,
         "start_offset": codec.hybrid_start_offset(l[0], 16), "message_bits": len(bits)}
st4, ln4, off4, tot4 = codec.extraction_plan(meta4, npx)
d_bits = torch.zeros((tot4 + 7) // 8 + 16, dtype=torch.uint8, device=dev)
t = dev_time(lambda: _cabi.check(L.peeb_lsb_extract(ws.handle, d_stego.data_ptr(), npx, 2, s, st4.ctypes.data, ln4.ctypes.data,
                                                   off4.ctypes.data, tot4, d_bits.data_ptr(), stream())))
t_api = wall(lambda: codec.extract_message_bits(stego_img, meta4), reps=1)
t0 = time.perf_counter(); r4 = OC.extract_message_bits(stego_img, meta4); t_cpu = time.perf_counter() - t0
assert r4 == codec.extract_message_bits(stego_img, meta4) == bits[:tot4]
rows.append({"row": "N4", "kernel": "lsb_extract_kernel", "workload": f"{tot4 / 1e6:.1f} Mbit read back from s={s} planes of 3000x3000 u16",
             "device_ms": t * 1e3, "mbit_s": tot4 / t / 1e6, "algorithmic_gb_s": (2 * tot4 + tot4 / 8) / t / 1e9,
             "frac_of_measured_peak": (2 * tot4 + tot4 / 8) / t / 1e9 / peak, "api_ms": t_api * 1e3, "cpu_oracle_ms": t_cpu * 1e3,
             "api": "extract_message_bits (equals the embedded bit string and the restatement)"})

# ---- a10 sweep (BASELINE configs[4]): every T = 1..64 on 2048x2048 16-bit images, (image, T) = unit
from codec_tcc_b200 import shard  # noqa: E402
from oracle import pee_c  # noqa: E402

n_sw, hs = 8, 2048
sw_imgs = np.stack([synth_image(hs, hs, 65535, 300 + k) for k in range(n_sw)])
sw_pays = np.random.default_rng(3).integers(0, 256, (n_sw, hs * hs // 8), dtype=np.uint8)
Ts = list(range(1, 65))
t_api = wall(lambda: shard.sweep_sharded(sw_imgs, sw_pays, Ts, 16), reps=2)
table = shard.sweep_sharded(sw_imgs, sw_pays, Ts, 16)
t0 = time.perf_counter()
for T in (1, 16, 64):
    _, _, i0 = pee_c.embed(sw_imgs[0], sw_pays[0], hs * hs, T, 16)
    row = table[T - 1]
    assert [int(v) for v in row[3:8]] == [i0[k] for k in ("capacity", "cap0", "cap1", "n_flagged", "sse")]
t_cpu = (time.perf_counter() - t0) / 3
units = n_sw * len(Ts)
rows.append({"row": "a10-sweep", "kernel": "pee_count + pee_embed (statistics only, shared cover and payload)",
             "workload": f"{n_sw} images {hs}x{hs} u16 x T=1..64 = {units} (image,T) units", "api_ms": t_api * 1e3,
             "api_mpixel_s": units * hs * hs / t_api / 1e6, "cpu_oracle_ms": t_cpu * 1e3 * units,
             "cpu_oracle_mpixel_s": hs * hs / t_cpu / 1e6,
             "api": "shard.sweep_sharded -> pee.pee_sweep_pairs (numpy in, table out); CPU = C oracle, 1 core, extrapolated from 3 units; 3 rows checked equal"})

print(json.dumps({"peak_gb_s": peak, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if os.path.exists(pp) else "fallback",
                  "cpu_cores_used_by_oracle": 1, "rows": rows}, indent=1))
