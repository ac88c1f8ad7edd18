#!/usr/bin/env python
"""Headline benchmark: PEE embed + extract round trip, Mpixel/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one max-capacity embed followed by one extract + recovery over one
batch of DICOM-shaped images.  Workloads (BASELINE.json configs):
  ct512     512 slices of 512x512, 16-bit  (configs[2]; the headline; PER GPU -> weak scaling)
  dx3000    64 radiographs of 3000x3000, 12-bit in uint16  (configs[3])
  slice     one 512x512 16-bit slice  (configs[1]; latency case, L2 flushed between steps)
  pe        the reference's own MR image images/pe.dcm (configs[0]; pixels from tests/golden/fixtures.npz,
            512 copies as one batch) + the reference's LSB flow of main() on it through embed_pipeline
  ct512sat  ct512 with large clipped regions and hard edges (the generic-code stress case)
  sweep2048 threshold sweep T = 1..64 over 64 images of 2048x2048, 16-bit (configs[4]; the (image, T)
            grid is sharded over the ranks; embed only: capacity / SSE table)
Printed: ONE JSON line (rank 0).  `value` is device-resident throughput (inputs
in HBM when the clock starts), `e2e` the same metric through the numpy API with
pinned host buffers and both PCIe copies inside the timed region; the other
workloads follow the headline timing as `other_workloads` (device resident).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (n_images per GPU, h, w, maxval, bit_depth, T)
    "ct512": (512, 512, 512, 65535, 16, 96),
    "dx3000": (64, 3000, 3000, 4095, 12, 12),
    "slice": (1, 512, 512, 65535, 16, 96),
    "pe": (512, 512, 512, 4095, 12, 16),
    "ct512sat": (512, 512, 512, 65535, 16, 96),
}
OTHER_WORKLOADS = ("slice", "dx3000", "pe", "ct512sat", "sweep2048")
SWEEP = {"n_images": 64, "h": 2048, "w": 2048, "bit_depth": 16, "T": tuple(range(1, 65)), "distinct": 16}
METRIC = "pee_embed_extract_roundtrip_throughput"
UNIT = "Mpixel/s"

# ------------------------------------------------------------------ CPU legs (oracle port of the numpy path)
_CPU = {}


def _cpu_roundtrip(k):
    """One image through the numpy oracle: embed at its capacity, extract, recover."""
    from oracle import pee_numpy as PN

    img, pay, T, bd = _CPU["imgs"][k], _CPU["pays"][k], _CPU["T"], _CPU["bd"]
    bits = PN.payload_to_bits(pay, None)
    maxval = (1 << bd) - 1
    cur, lm, cap0, cap1 = PN.embed_fixed_T(img, bits, T, maxval)  # zero-padded: every carrier is expanded
    cap = cap0 + cap1
    marked = cur.astype(img.dtype)
    out, rec = PN.pee_extract(marked, np.packbits(lm, axis=1), T, cap, bd)
    ok = bool(np.array_equal(rec, img)) and bool(np.array_equal(np.unpackbits(out)[:cap], bits[:cap]))
    return img.size, ok


def cpu_port_throughput(name, n_sample, repeats=1):
    """Mpixel/s of the numpy oracle over `n_sample` images of the workload, image
    parallel over every host core this process may use (fork pool)."""
    import multiprocessing as mp

    from codec_tcc_b200.synth import synth_image

    n, h, w, maxval, bd, T = WORKLOADS[name]
    rng = np.random.default_rng(99)
    _CPU.update(T=T, bd=bd)
    _CPU["imgs"] = [synth_image(h, w, maxval, 1000 + k) for k in range(n_sample)]
    _CPU["pays"] = [rng.integers(0, 256, (h * w + 7) // 8, dtype=np.uint8) for _ in range(n_sample)]
    cores = len(os.sched_getaffinity(0))
    procs = max(1, min(cores, n_sample))
    times = []
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_cpu_roundtrip, range(min(procs, n_sample)))  # warm the workers
        for _ in range(repeats):
            t0 = time.perf_counter()
            res = pool.map(_cpu_roundtrip, range(n_sample), chunksize=1)
            times.append(time.perf_counter() - t0)
    assert all(ok for _, ok in res), "CPU oracle round trip failed"
    px = sum(p for p, _ in res)
    return px, times, procs


def cpu_c_throughput(name, n_sample):
    """The scalar C restatement (oracle/pee_ref.c, OpenMP over images) on the same workload: what a compiled CPU
    implementation of the same specification does on this box's cores -- context for the numpy figure."""
    from codec_tcc_b200.synth import synth_batch
    from oracle import pee_c

    n, h, w, maxval, bd, T = WORKLOADS[name]
    imgs = synth_batch(n_sample, h, w, maxval, 1000)
    stride = (h * w + 7) // 8 + 8
    pays = np.random.default_rng(99).integers(0, 256, (n_sample, stride), dtype=np.uint8)
    _, _, info = pee_c.embed_batch(imgs, pays, np.full(n_sample, h * w, np.int64), T, bd)  # every carrier takes a payload bit
    cap = info[:, 2].copy()
    t0 = time.perf_counter()
    marked, lm, info = pee_c.embed_batch(imgs, pays, cap, T, bd)
    out, rec, rc = pee_c.extract_batch(marked, lm, T, cap, stride)
    sec = time.perf_counter() - t0
    assert rc == 0 and np.array_equal(rec, imgs)
    return {"value": n_sample * h * w / sec / 1e6, "unit": UNIT, "cores": int(pee_c.threads()),
            "what": f"oracle/pee_ref.c embed + extract of {n_sample} images, OpenMP over images"}


def run_reference_arm(args):
    """--impl reference: the reference is pure numpy and ships no PEE code
    (SURVEY.md F2), so the CPU arm is the numpy oracle port of the same path on
    the host cores, a bounded sample of the workload per step."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    n, h, w, maxval, bd, T = WORKLOADS[args.workload]
    per_step = max(1, min(n, 4 * cores if h * w <= 1 << 20 else max(1, cores // 2)))
    px, times, procs = cpu_port_throughput(args.workload, per_step, repeats=args.steps + args.warmup)
    timed = times[args.warmup:]
    sec = sum(timed) / len(timed)
    value = px / sec / 1e6
    sample = f"{per_step} of {n} images of workload {args.workload} per step, numpy oracle, fork pool of {procs}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16" if maxval > 255 else "u8", "data": "synthetic",
        "config": workload_config(args.workload, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(name, world, distinct=None):
    n, h, w, maxval, bd, T = WORKLOADS[name]
    kind = {"pe": "copies of the reference's images/pe.dcm", "ct512sat": "synthetic images with clipped regions and hard edges"}.get(name, "synthetic images")
    rep = f" ({distinct} distinct)" if distinct and distinct != n else ""
    return {
        "workload": f"{name}: {n} {kind}/GPU of {h}x{w}{rep}, {bd}-bit in uint{16 if maxval > 255 else 8}, "
                    f"PEE T={T}, payload = capacity (max-capacity embed + extract + recovery)",
        "images_per_gpu": n, "height": h, "width": w, "bit_depth": bd, "T": T,
        "parallelism": f"image-sharded x{world}, no data-path collective",
        "l2": "inputs larger than L2 (no flush needed)" if n * h * w * 2 > 130e6 else "L2 flushed between steps",
    }


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons of one GPU, sampled DURING the timed region.  NVML in process
    (initialised before the region, a few microseconds per sample); starting an `nvidia-smi -lms`
    child instead puts its start-up (driver enumeration of every GPU of the box) inside the region,
    which showed up once as a straggling rank of a multi-GPU run.  nvidia-smi stays as the fallback."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"),
               (0x80, "hw_power_brake_slowdown"))

    def __init__(self, gpu_index, uuid=None):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.nvml = self.handle = None
        self.samples, self.bits, self.stop_flag = [], 0, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = (pynvml.nvmlDeviceGetHandleByUUID(uuid if isinstance(uuid, bytes) else uuid.encode())
                           if uuid else pynvml.nvmlDeviceGetHandleByIndex(gpu_index))
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:  # noqa: BLE001 - no NVML binding / no permission: nvidia-smi child below
            self.nvml = None

    def _poll(self):
        nv, hd = self.nvml, self.handle
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM)))
                self.bits |= int(reasons(hd))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.001)

    def sample_now(self):
        """A few samples taken from the calling thread (used while the timed steps are still queued on the GPU)."""
        if self.nvml is None:
            return
        nv, hd = self.nvml, self.handle
        reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        for _ in range(3):
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(hd, nv.NVML_CLOCK_SM)))
                self.bits |= int(reasons(hd))
            except Exception:  # noqa: BLE001
                pass

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=1.0)
            sm = self.samples
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax,
                    "reasons": sorted(name for bit, name in self.REASONS if self.bits & bit), "samples": len(sm),
                    "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------ the GPU arm
def workload_images(name, n, h, w, maxval, seed):
    """Host images of a workload: synthetic (SURVEY.md 8d generators) or the reference's own fixture.
    Large synthetic batches repeat a few distinct images (their generation on the host would take longer
    than the whole measurement); the distinct count is part of the workload description."""
    from codec_tcc_b200.synth import synth_batch, synth_saturated

    dtype = np.uint16 if maxval > 255 else np.uint8
    if name == "pe":
        z = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
        return np.broadcast_to(z["pe"].astype(dtype), (n, h, w)), 1
    distinct = n if h * w <= 1 << 20 else min(n, 8)
    if name == "ct512sat":
        base = np.stack([synth_saturated(h, w, maxval, seed + k) for k in range(distinct)]).astype(dtype)
    else:
        base = synth_batch(distinct, h, w, maxval, seed)
    if distinct == n:
        return base, distinct
    return base[np.arange(n) % distinct], distinct


class RoundTrip:
    """Device-resident embed + extract of one workload on the current device (the timed unit of bench.py)."""

    def __init__(self, name, rank, dev, dims=None):
        import torch

        from codec_tcc_b200 import _cabi, device as D

        self.name, self.dev, self.D, self.torch = name, dev, D, torch
        n, h, w, maxval, bd, T = dims or WORKLOADS[name]
        self.n, self.h, self.w, self.maxval, self.bd, self.T = n, h, w, maxval, bd, T
        self.npx = n * h * w
        t0 = time.perf_counter()
        imgs, self.distinct = workload_images(name, n, h, w, maxval, 2 + rank * n)
        self.imgs_h = _cabi.pinned_empty((n, h, w), imgs.dtype)
        self.imgs_h[...] = imgs
        self.stride = D.payload_stride(h * w)
        self.pays_h = _cabi.pinned_empty((n, self.stride), np.uint8)
        self.pays_h[...] = np.random.default_rng(7 + rank).integers(0, 256, (n, self.stride), dtype=np.uint8)
        ih = self.imgs_h
        self.d_imgs = torch.from_numpy(ih.view(np.int16) if ih.dtype == np.uint16 else ih).to(dev)
        self.d_pays = torch.from_numpy(self.pays_h).to(dev)
        self.gen_s = time.perf_counter() - t0
        # capacity of every image with this payload stream (one untimed embed, nothing written)
        big = np.full(n, h * w, np.int64)
        _, _, d_info = D.pee_embed_device(self.d_imgs, self.d_pays, big, T, bd, marked=False, lm=False)
        self.cap = d_info[:, 2].cpu().numpy().astype(np.int64)
        assert (self.cap > 0).all()
        self.d_marked = torch.empty_like(self.d_imgs)
        self.d_lm = torch.empty((n, h, (w + 7) // 8), dtype=torch.uint8, device=dev)
        self.d_rec = torch.empty_like(self.d_imgs)
        self.d_out = torch.empty((n, self.stride), dtype=torch.uint8, device=dev)
        self.d_info_e = torch.empty((n, 8), dtype=torch.int64, device=dev)
        self.d_info_x = torch.empty((n, 8), dtype=torch.int64, device=dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if self.npx * ih.dtype.itemsize <= 130e6 else None

    def step(self):
        D = self.D
        D.pee_embed_device(self.d_imgs, self.d_pays, self.cap, self.T, self.bd, marked=self.d_marked, lm=self.d_lm,
                           info=self.d_info_e)
        D.pee_extract_device(self.d_marked, self.d_lm, self.T, self.cap, self.bd, payload_out=self.d_out,
                             recovered=self.d_rec, info=self.d_info_x)

    def timed_steps(self, k, while_busy=None):
        """K steps, CUDA events on the launching stream; with a small workload the L2 is flushed between steps
        and only the steps are timed.  -> total ms."""
        torch = self.torch
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        for a, b in ev:
            if self.flush is not None:
                self.flush.fill_(1)
            a.record()
            self.step()
            b.record()
        if while_busy is not None:  # the host is ahead of the GPU here: the queued steps are still running
            while_busy()
        torch.cuda.synchronize(self.dev)
        return sum(a.elapsed_time(b) for a, b in ev)

    def check(self, oracle_units):
        """What was just timed: identity round trip on every image, payload on a strided sample, and marked image /
        location map / statistics of `oracle_units` strided images against the CPU oracle (oracle/pee_ref.c)."""
        torch = self.torch
        n = self.n
        assert torch.equal(self.d_rec, self.d_imgs), "recovered images differ from the originals"
        xi, ei = self.d_info_x.cpu().numpy(), self.d_info_e.cpu().numpy()
        assert (xi[:, 7] == 0).all() and (ei[:, 7] == 0).all()
        assert np.array_equal(xi[:, 2], self.cap) and np.array_equal(ei[:, 2], self.cap)
        out_h = self.d_out.cpu().numpy()
        for u in range(0, n, max(1, n // 64)):
            nb_, rem = int(self.cap[u]) // 8, int(self.cap[u]) % 8
            assert np.array_equal(out_h[u, :nb_], self.pays_h[u, :nb_]), "extracted payload differs"
            if rem:
                assert int(out_h[u, nb_]) == int(self.pays_h[u, nb_]) & ((0xFF00 >> rem) & 0xFF)
        checked = 0
        if oracle_units:
            from oracle import pee_c
            units = sorted(set(range(0, n, max(1, n // oracle_units))))[:oracle_units]
            sel = torch.as_tensor(units, device=self.dev)
            marked_h = self.d_marked[sel].cpu().numpy().view(self.imgs_h.dtype)
            lm_h = self.d_lm[sel].cpu().numpy()
            # (the C restatement over all selected units at once, one OpenMP thread per image)
            stride = self.pays_h.shape[1]
            pays = np.zeros((len(units), stride + 8), np.uint8)
            pays[:, :stride] = self.pays_h[units]
            m0, lm0, i0 = pee_c.embed_batch(self.imgs_h[units], pays, self.cap[units], self.T, self.bd)
            bad = np.flatnonzero((marked_h != m0).reshape(len(units), -1).any(axis=1) | (lm_h != lm0).reshape(len(units), -1).any(axis=1))
            assert bad.size == 0, f"units {[units[k] for k in bad[:8]]}: GPU embed differs from the CPU oracle"
            assert np.array_equal(i0[:, 2:7], ei[units][:, 2:7]), "capacity / flagged / SSE differ from the CPU oracle"
            checked = len(units)
        return checked

    def kernel_times(self, steps):
        """Per-kernel device time (events around every launch; separate pass, not the timed one) and the share of
        warp-steps of the embed kernel that took the generic code."""
        from codec_tcc_b200 import _cabi

        ws = _cabi.workspace(self.dev.index)
        ws.prof_enable(True)
        ws.step_counters(True)
        for _ in range(steps):
            if self.flush is not None:
                self.flush.fill_(1)
            self.step()
        self.torch.cuda.synchronize(self.dev)
        prof = ws.prof_report()
        ws.prof_enable(False)
        sc = ws.step_counters(False)
        item = self.imgs_h.dtype.itemsize
        alg = {"pee_embed": 2 * item + 0.25, "pee_extract": 2 * item + 0.25, "pee_count": float(item),
               "pee_gather": 0.25, "pee_finalize": 0.0}
        kernels = {}
        for kname, (ms, calls) in prof.items():
            avg = ms / calls
            kernels[kname] = {"avg_ms": avg, "launches_per_step": calls / steps,
                              "algorithmic_gb_per_s": (alg.get(kname, 0.0) * self.npx / 1e9) / (avg * 1e-3) if avg > 0 else None}
        tot = max(sc["steps"], 1)
        generic = {"embed_warp_steps": sc["steps"], "at_image_border": sc["edge"] / tot, "redone_out_of_range": sc["redone"] / tot}
        return kernels, alg, generic


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def other_roundtrip(name, rank, dev, steps=5):
    """One of the non-headline round-trip workloads, device resident: throughput, fraction of the HBM roofline of the
    whole step, kernel times, generic-code share, oracle check on a strided sample."""
    import torch

    rt = RoundTrip(name, rank, dev)
    # warm up for at least ~40 ms of GPU work: the host-side set-up of a workload lets the clocks fall back
    t0 = time.perf_counter()
    while True:
        for _ in range(3):
            rt.step()
        torch.cuda.synchronize(dev)
        if time.perf_counter() - t0 > 0.04:
            break
    if rt.flush is not None:
        rt.timed_steps(steps)  # (small workloads: one untimed window with the L2 flush in the loop; the first one measured 80-90 us
                               # against 51 us for every later one, scripts/slice_latency.py)
    ms = rt.timed_steps(steps) / steps
    checked = rt.check(4 if rt.h * rt.w > 1 << 20 else 16)
    kernels, _, generic = rt.kernel_times(3)
    item = rt.imgs_h.dtype.itemsize
    peak, _ = hbm_peak()
    gbs = (4 * item + 0.5) * rt.npx / 1e9 / (ms * 1e-3)
    out = {"config": workload_config(name, 1, rt.distinct), "value": rt.npx / (ms * 1e-3) / 1e6, "unit": UNIT,
           "ms_per_step": ms, "steps": steps, "step_algorithmic_gb_per_s": gbs, "frac": gbs / peak,
           "capacity_bpp": float(rt.cap.mean() / (rt.h * rt.w)),
           "kernel_ms": {k: round(v["avg_ms"], 5) for k, v in kernels.items()}, "generic_code": generic,
           "oracle_units_checked": checked}
    if name == "slice":
        out["latency_us_per_round_trip"] = ms * 1e3
        out["path"] = "cluster kernels (one launch per direction)" if "pee_count" not in kernels else "band kernels"
        # the same slice through the band kernels (count, embed, extract, gather), for comparison
        prev = os.environ.get("PEEB_CLUSTER")
        os.environ["PEEB_CLUSTER"] = "0"
        try:
            for _ in range(3):
                rt.step()
            out["latency_us_band_kernels"] = rt.timed_steps(steps) / steps * 1e3
        finally:
            if prev is None:
                os.environ.pop("PEEB_CLUSTER", None)
            else:
                os.environ["PEEB_CLUSTER"] = prev
    return out


def pe_lsb_flow(dev):
    """configs[0], the part the reference really implements: steps 3-5 of its main() (src/codec.py:868-880:
    decomposition at beta = 0.4, hybrid LSB embed with 16x16 search tiles, merge) on images/pe.dcm with main()'s
    message, through codec_tcc_b200.codec.embed_pipeline (numpy in / numpy out), beside the numpy restatement of the
    same three calls (oracle/codec_numpy.py, pinned byte for byte to the reference's outputs on this image by
    tests/golden) on one host core."""
    from codec_tcc_b200 import codec
    from oracle import codec_numpy as OC

    img = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))["pe"]
    bits = codec.message_to_bits("Mensagem de teste para esteganografia!")

    def gpu():
        return codec.embed_pipeline(img, bits, beta=0.4, search_block_size=16, device=dev.index)

    def cpu():
        g, loc = OC.adaptive_modalities_decomposition(img, 0.4)
        st, bm, used, lens, idx = OC.lsb_embed_block_then_multiplane(loc, bits, search_block_size=16)
        return OC.merge_modalities(g, st), bm

    stego, bitmaps, meta = gpu()
    ref_stego, ref_bm = cpu()
    assert np.array_equal(stego, ref_stego) and np.array_equal(bitmaps, np.stack(ref_bm)), "LSB flow differs from the restatement"

    def wall(fn, reps):
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    gpu_s, cpu_s = wall(gpu, 10), wall(cpu, 3)
    return {"what": "reference main() steps 3-5 on images/pe.dcm (512x512, 12 bit), 304-bit message, numpy in / numpy out",
            "gpu_ms_per_image": gpu_s * 1e3, "cpu_ms_per_image": cpu_s * 1e3, "cpu_kind": "port",
            "cpu_note": "numpy restatement pinned to the reference's outputs on this image (tests/golden), 1 core",
            "bit_exact": True, "s": int(meta["s"])}


def sweep2048(rank, world, dev, barrier, all_max):
    """configs[4]: every threshold T = 1..64 on every image of a series of 2048x2048 16-bit images; the unit of work is
    the (image, T) pair, contiguous blocks of the flat grid per rank (shard.partition_grid), no exchange; one real
    embed per pair (pass 1 depends on pass 0's output), statistics only (capacity, SSE -> PSNR)."""
    import torch

    from codec_tcc_b200 import device as D, shard
    from codec_tcc_b200.synth import synth_batch

    n_img, h, w, bd = SWEEP["n_images"], SWEEP["h"], SWEEP["w"], SWEEP["bit_depth"]
    Ts = np.asarray(SWEEP["T"], np.int32)
    img_idx, t_idx = shard.partition_grid(n_img, Ts.size, world, rank)
    distinct = SWEEP["distinct"]
    mine = sorted(set(int(i) for i in img_idx))
    base = synth_batch(distinct, h, w, (1 << bd) - 1, 5)
    stride = D.payload_stride(h * w)
    pay = torch.from_numpy(np.random.default_rng(11).integers(0, 256, (1, stride), dtype=np.uint8)).to(dev).repeat(Ts.size, 1)
    d_imgs = {i: torch.from_numpy(base[i % distinct].view(np.int16)).to(dev) for i in mine}
    groups = [(i, Ts[t_idx[img_idx == i]]) for i in mine]
    infos = [torch.empty((len(t), 8), dtype=torch.int64, device=dev) for _, t in groups]

    def run():
        for (i, t), info in zip(groups, infos):
            D.pee_embed_device(d_imgs[i], pay[:len(t)], np.full(len(t), h * w, np.int64), t, bd, marked=False, lm=False,
                               info=info, shared_cover=True)

    run()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run()
    b.record()
    torch.cuda.synchronize(dev)
    barrier()
    ms = all_max(a.elapsed_time(b))
    pairs = n_img * Ts.size
    # sanity of the table: capacity grows with T, SSE too
    first = infos[0].cpu().numpy()
    assert (np.diff(first[:, 2]) >= 0).all() and (first[:, 7] <= 0).all()
    peak, _ = hbm_peak()
    gbs = 2.0 * pairs * h * w / 1e9 / (ms * 1e-3)  # statistics only: the image is read once per pair, nothing written
    return {"config": {"workload": f"sweep2048: T = 1..{Ts.size} on {n_img} images of {h}x{w}, {bd}-bit ({distinct} distinct), "
                                   f"{pairs} (image, T) embeds, statistics only", "parallelism": f"(image, T) grid sharded x{world}"},
            "value": pairs * h * w / (ms * 1e-3) / 1e6, "unit": "Mpixel/s embedded", "ms": ms, "pairs": pairs,
            "algorithmic_gb_per_s": gbs, "frac": gbs / peak / world,
            "table_head": [{"T": int(r[0]), "capacity": int(r[2]), "sse": int(r[6])} for r in first[:4]]}


def auto_threshold(dev):
    """Threshold selection for a batch on the device (pee_embed_device(T=None), Appendix A): error histogram ->
    estimate per image -> embed -> only the images that fall short are embedded again at T + 1.  64 slices of
    512x512, 16 bit, payloads of 20-90 % of the capacity at T = 96; the chosen T and the marked images of a strided
    sample against the CPU oracle's own search."""
    import torch

    from codec_tcc_b200 import device as D
    from codec_tcc_b200.synth import random_payload, synth_batch
    from oracle import pee_numpy as PN

    n, h, w, bd = 64, 512, 512, 16
    imgs = synth_batch(n, h, w, 65535, 21)
    d_imgs = torch.from_numpy(imgs.view(np.int16)).to(dev)
    stride = D.payload_stride(h * w)
    pays = np.random.default_rng(3).integers(0, 256, (n, stride), dtype=np.uint8)
    d_pays = torch.from_numpy(pays).to(dev)
    _, _, info = D.pee_embed_device(d_imgs, d_pays, np.zeros(n, np.int64), 96, bd, marked=False, lm=False)
    cap96 = info[:, 2].cpu().numpy()
    nb = (cap96 * np.linspace(0.2, 0.9, n)).astype(np.int64)
    d_marked = torch.empty_like(d_imgs)
    d_lm = torch.empty((n, h, (w + 7) // 8), dtype=torch.uint8, device=dev)

    def run():
        return D.pee_embed_device(d_imgs, d_pays, nb, None, bd, marked=d_marked, lm=d_lm)

    run()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        _, _, info = run()
    torch.cuda.synchronize(dev)
    ms = (time.perf_counter() - t0) / reps * 1e3
    info = info.cpu().numpy()
    assert (info[:, 7] == 0).all() and (info[:, 2] >= nb).all()
    marked = d_marked.cpu().numpy().view(np.uint16)
    checked = 0
    for u in range(0, n, 16):
        m0, _, i0 = PN.pee_embed(imgs[u], pays[u], None, bd, n_bits=int(nb[u]))
        assert int(info[u, 0]) == i0["T"] and np.array_equal(marked[u], m0), f"threshold selection differs from the oracle on image {u}"
        checked += 1
    fixed = time.perf_counter()
    for _ in range(reps):
        D.pee_embed_device(d_imgs, d_pays, nb, info[:, 0].astype(np.int32), bd, marked=d_marked, lm=d_lm)
    torch.cuda.synchronize(dev)
    fixed_ms = (time.perf_counter() - fixed) / reps * 1e3
    return {"what": f"{n} slices of {h}x{w}, {bd}-bit: smallest T per image that holds its payload (20-90 % of the capacity at T = 96), chosen on the device",
            "ms_per_batch": ms, "value": n * h * w / (ms * 1e-3) / 1e6, "unit": "Mpixel/s embedded (wall clock, search included)",
            "ms_embed_at_known_T": fixed_ms, "T_chosen_min_max": [int(info[:, 0].min()), int(info[:, 0].max())],
            "oracle_units_checked": checked}


def bitmap_coding(dev):
    """N2: the side bitmaps of the LSB embedders as a "PBR1" blob (bit packing + zero-run elimination on the device)
    beside the reference's blob step, zlib over one byte per pixel (src/codec.py:888-889), on the host.  Input: the
    bitmaps of four 3000x3000 images with s = 4 planes each (144 MB), with about 2 M (sparse) and 36 M (dense)
    changed positions; device resident, CUDA events; the blob decodes back to the input."""
    import zlib

    import torch

    from codec_tcc_b200 import device as D
    from oracle import bitcode_numpy as BN

    from codec_tcc_b200 import _cabi

    h = w = 3000
    s = 16          # four images' worth of s = 4 bitmaps: 144 MB, larger than L2
    rng = np.random.default_rng(8)
    ws = _cabi.workspace(dev.index)
    out = {}
    peak, _ = hbm_peak()
    for label, bits in (("sparse_4Mbit", 4 << 20), ("dense_72Mbit", 72 << 20)):
        maps = np.zeros(s * h * w, np.uint8)
        pos = rng.choice(maps.size, size=bits // 2, replace=False)   # about half of the written positions flip
        maps[pos] = 1
        d_maps = torch.from_numpy(maps).to(dev)
        blob = D.bitmap_encode_device(d_maps)
        back = D.bitmap_decode_device(blob, maps.size)
        assert torch.equal(back, d_maps), "PBR1 round trip is not the identity"
        if label == "sparse_4Mbit":
            assert bytes(blob.cpu().numpy()) == BN.encode(maps), "PBR1 blob differs from the CPU restatement"
        reps = 5
        full = torch.empty(int(_cabi_bound(maps.size)), dtype=torch.uint8, device=dev)
        # kernel time from the library's own CUDA events around its launches (each call also synchronises once to
        # hand the blob size / the verdict on the blob back to the host: that wait is in the wall-clock figures)
        ws.prof_enable(True)
        t0 = time.perf_counter()
        for _ in range(reps):
            blob = D.bitmap_encode_device(d_maps, blob=full)
        t1 = time.perf_counter()
        for _ in range(reps):
            D.bitmap_decode_device(blob, maps.size, out=back)
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        prof = ws.prof_report()
        ws.prof_enable(False)
        enc_ms, dec_ms = prof["bitmap_encode"][0] / prof["bitmap_encode"][1], prof["bitmap_decode"][0] / prof["bitmap_decode"][1]
        enc_wall, dec_wall = (t1 - t0) / reps * 1e3, (t2 - t1) / reps * 1e3
        sample = maps[: maps.size // 8]
        t0 = time.perf_counter()
        z = zlib.compress(sample.tobytes())
        z_s = (time.perf_counter() - t0) * 8
        out[label] = {"elements": int(maps.size), "blob_bytes": int(blob.numel()), "zlib_bytes_estimate": len(z) * 8,
                      "encode_ms": enc_ms, "decode_ms": dec_ms, "encode_wall_ms": enc_wall, "decode_wall_ms": dec_wall,
                      "encode_gb_per_s": maps.size / 1e9 / (enc_ms * 1e-3), "encode_frac_of_hbm": maps.size / 1e9 / (enc_ms * 1e-3) / peak,
                      "decode_gb_per_s": maps.size / 1e9 / (dec_ms * 1e-3),
                      "cpu_zlib_ms": z_s * 1e3, "cpu_note": "zlib.compress of one eighth of the bytes, scaled by 8, 1 core (the reference's blob step)"}
    return out


def _cabi_bound(n):
    from codec_tcc_b200 import _cabi
    return _cabi.lib().peeb_bitmap_blob_bound(int(n))


def run_gpu_arm(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    name = args.workload
    n, h, w, maxval, bd, T = WORKLOADS[name]

    # CPU baseline first (fork pool before CUDA is initialised in this process)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        big = h * w > 1 << 20
        n_sample = 12 if big else 192
        px, times, procs = cpu_port_throughput(name if name in ("ct512", "dx3000", "slice") else "ct512", n_sample)
        cpu_baseline = {"value": px / times[0] / 1e6, "unit": UNIT, "cores": procs, "kind": "port",
                        "sample": f"{n_sample} images of {h}x{w} from the same generator, numpy oracle "
                                  f"(oracle/pee_numpy.py) embed+extract, fork pool of {procs}, {times[0]:.1f} s wall"}
        try:
            cpu_baseline["c_restatement"] = cpu_c_throughput(name if name in ("ct512", "dx3000", "slice") else "ct512",
                                                             4 if big else 128)
        except Exception as exc:  # noqa: BLE001
            cpu_baseline["c_restatement"] = {"error": f"{type(exc).__name__}: {exc}"}

    import torch
    import torch.distributed as dist

    from codec_tcc_b200 import _cabi, pee, shard

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = shard.bind_to_gpu_numa(local) if world > 1 else []  # host buffers on the GPU's own NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    os.environ["PEEB_DEVICE"] = str(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def all_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    rt = RoundTrip(name, rank, dev)
    npx = rt.npx
    for _ in range(args.warmup):
        rt.step()
    barrier()
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:  # noqa: BLE001
        gpu_uuid = None
    sampler = ClockSampler(local, gpu_uuid)
    if rank == 0:
        sampler.start()
    barrier()
    ms_total = rt.timed_steps(args.steps, sampler.sample_now if rank == 0 else None)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    rank_ms = [ms_total / args.steps]
    if world > 1:
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        rank_ms = [float(x.item()) / args.steps for x in every]  # reported beside the max: a straggler shows
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = npx * world / (ms_step * 1e-3) / 1e6

    # ---- correctness of what was just timed: identity round trip on all images, payloads, and a strided sample of
    # every image of rank 0's shard (marked image, location map, statistics) against the CPU oracle -- outside the timed region
    oracle_checked = rt.check(n if rank == 0 else 0)  # every image of rank 0's shard

    kernels, alg, generic = rt.kernel_times(args.steps)
    item = rt.imgs_h.dtype.itemsize
    peak, peak_src = hbm_peak()
    dom = max(kernels, key=lambda k: kernels[k]["avg_ms"] * kernels[k]["launches_per_step"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(name, {}).get(dom)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["algorithmic_gb_per_s"], "peak": peak,
                "unit": "GB/s", "frac": kernels[dom]["algorithmic_gb_per_s"] / peak, "traffic": traffic,
                "peak_source": peak_src,
                "algorithmic_bytes_per_pixel": alg[dom],
                "step_algorithmic_gb_per_s": (4 * item + 0.5) * npx / 1e9 / (ms_step * 1e-3),
                "step_frac": (4 * item + 0.5) * npx / 1e9 / (ms_step * 1e-3) / peak}

    # ---- end to end through the numpy API: pinned host buffers, both copies inside the clock
    imgs_h, pays_h, cap = rt.imgs_h, rt.pays_h, rt.cap
    marked_p = _cabi.pinned_empty((n, h, w), imgs_h.dtype)
    lm_p = _cabi.pinned_empty((n, h, (w + 7) // 8), np.uint8)
    rec_p = _cabi.pinned_empty((n, h, w), imgs_h.dtype)
    out_p = _cabi.pinned_empty((n, (h * w + 7) // 8), np.uint8)
    pay_p = _cabi.pinned_empty((n, (h * w + 7) // 8), np.uint8)
    pay_p[...] = pays_h[:, :pay_p.shape[1]]

    def e2e_step():
        _, _, ie = pee.pee_embed_batch(imgs_h, pay_p, cap, T, bd, out_marked=marked_p, out_lm=lm_p, device=local)
        _, _, ix = pee.pee_extract_batch(marked_p, lm_p, T, cap, bd, out_recovered=rec_p, out_payload=out_p, device=local)
        return ie, ix

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ie, ix = e2e_step()
    sec = time.perf_counter() - t0
    barrier()
    e2e_ms = all_max(sec) / e2e_steps * 1e3
    assert np.array_equal(rec_p, imgs_h) and (ie[:, 7] == 0).all() and (ix[:, 7] == 0).all()
    pcie = e2e_bytes(n, h, w, item, cap, ie)
    e2e = {"value": npx * world / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(pcie["h2d"]),
           "d2h_bytes_per_step": int(pcie["d2h"]), "ms_per_step": e2e_ms, "steps": e2e_steps,
           "api": "codec_tcc_b200.pee.pee_embed_batch + pee_extract_batch (numpy in / numpy out, pinned host buffers)"}
    floor = pcie_floor(world)
    if floor:
        # both directions run at once: the floor of a step is the larger direction at the per-direction rate
        e2e["pcie_floor_ms"] = max(pcie["h2d"], pcie["d2h"]) / (floor["per_rank_gbs_both_directions"] * 1e9) * 1e3
        e2e["frac_of_pcie_floor"] = e2e["pcie_floor_ms"] / e2e_ms
        e2e["pcie_floor_source"] = floor["source"]

    stats = shard.gather_stats(rt.d_info_e[:, :7]) if world > 1 else rt.d_info_e
    images_total = int(stats.shape[0])
    gen_s, capacity_bpp = rt.gen_s, float(cap.mean() / (h * w))
    del rt, marked_p, lm_p, rec_p, out_p, pay_p
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs and data distributions, after the headline timing
    others = {}
    if not args.no_other_workloads:
        for wl in OTHER_WORKLOADS:
            if wl == name:
                continue
            try:
                if wl == "sweep2048":
                    others[wl] = sweep2048(rank, world, dev, barrier, all_max)
                elif world == 1:
                    others[wl] = other_roundtrip(wl, rank, dev)
                    if wl == "pe":
                        others[wl]["lsb_flow"] = pe_lsb_flow(dev)
            except Exception as exc:  # noqa: BLE001 -- reported, never hides the headline
                others[wl] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()
        if world == 1:
            for key, fn in (("auto_threshold", auto_threshold), ("bitmap_coding", bitmap_coding)):
                try:
                    others[key] = fn(dev)
                except Exception as exc:  # noqa: BLE001
                    others[key] = {"error": f"{type(exc).__name__}: {exc}"}
                torch.cuda.empty_cache()

    if rank == 0:
        launches_per_step = sum(v["launches_per_step"] for v in kernels.values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16" if item == 2 else "u8", "data": "synthetic",
            "config": workload_config(name, world), "roofline": roofline, "kernels": kernels, "e2e": e2e,
            "gpu_launches": int(round(launches_per_step * args.steps)), "clocks": clocks,
            "capacity_bpp": capacity_bpp, "images_total": images_total,
            "bit_exact": f"round trip identity on all images; marked image, location map and statistics of {oracle_checked} "
                         "images (all of rank 0's) == CPU oracle (oracle/pee_ref.c); PEE parity is unpinned (no reference PEE exists)",
            "generic_code": generic,
            "numa_cores_rank0": len(numa_cores), "ms_per_step_by_rank": [round(x, 4) for x in rank_ms],
            "setup_s": gen_s, "other_workloads": others,
        }
        if cpu_baseline is not None:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_bytes(n, h, w, item, cap, info_embed):
    """Bytes one e2e step moves over PCIe in each direction (what the host wrappers copy: payload rows are moved up
    to the longest payload of the batch, rounded up to 16 bytes, not up to their stride)."""
    img = n * h * w * item
    lmb = n * h * ((w + 7) // 8)
    pay = n * ((int(((cap + 7) // 8).max()) + 15) // 16 * 16)
    h2d = img + pay + img + lmb                     # embed: images + payload bytes; extract: marked images + maps
    d2h = img + lmb + img + pay + 2 * n * 64        # embed: marked + maps; extract: recovered + payloads; info rows
    return {"h2d": h2d, "d2h": d2h}


def pcie_floor(world):
    """The measured host<->device copy floor of this pool's boxes with `world` ranks copying at once
    (profiles/pcie_floor_r02.json, written by scripts/pcie_probe.py under the same torchrun launch)."""
    path = os.path.join(ROOT, "profiles", "pcie_floor_r02.json")
    if not os.path.exists(path):
        return None
    rows = json.load(open(path)).get("by_world", {})
    row = rows.get(str(world))
    if not row:
        return None
    return {"per_rank_gbs_both_directions": row["per_rank_gbs_both_directions_per_direction"],
            "source": "profiles/pcie_floor_r02.json (scripts/pcie_probe.py, all ranks copying both ways at once)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="ct512", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-workloads", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
