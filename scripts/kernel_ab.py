#!/usr/bin/env python
"""Development aid: device-resident step time and per-kernel times of the PEE round trip for one workload
(the timed part of bench.py without its CPU legs), to compare library variants quickly.
    python scripts/kernel_ab.py [ct512|dx3000|slice|custom:n,h,w,bit_depth,T] [steps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, workload_images  # noqa: E402
from codec_tcc_b200 import _cabi, device as D  # noqa: E402
from codec_tcc_b200.synth import synth_batch  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ct512"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
if name.startswith("custom:"):  # custom:n,h,w,bit_depth,T
    n, h, w, bd, T = (int(x) for x in name[7:].split(","))
    maxval = (1 << bd) - 1
else:
    n, h, w, maxval, bd, T = WORKLOADS[name]
dev = torch.device("cuda:0")
imgs = synth_batch(n, h, w, maxval, 2) if name.startswith("custom:") or name in ("ct512", "dx3000", "slice") else workload_images(name, n, h, w, maxval, 2)[0]
d_imgs = torch.from_numpy(imgs.view(np.int16) if imgs.dtype == np.uint16 else imgs).to(dev)
stride = D.payload_stride(h * w)
d_pays = torch.from_numpy(np.random.default_rng(7).integers(0, 256, (n, stride), dtype=np.uint8)).to(dev)
_, _, d_info = D.pee_embed_device(d_imgs, d_pays, np.full(n, h * w, np.int64), T, bd, marked=False, lm=False)
cap = d_info[:, 2].cpu().numpy().astype(np.int64)
d_marked = torch.empty_like(d_imgs); d_lm = torch.empty((n, h, (w + 7) // 8), dtype=torch.uint8, device=dev)
d_rec = torch.empty_like(d_imgs); d_out = torch.empty((n, stride), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if n * h * w * 2 <= 130e6 else None


def step():
    D.pee_embed_device(d_imgs, d_pays, cap, T, bd, marked=d_marked, lm=d_lm)
    D.pee_extract_device(d_marked, d_lm, T, cap, bd, payload_out=d_out, recovered=d_rec)


for _ in range(3):
    step()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
for a, b in ev:
    if flush is not None:
        flush.fill_(1)
    a.record(); step(); b.record()
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ev)
assert torch.equal(d_rec, d_imgs)
ws = _cabi.workspace(0)
ws.prof_enable(True)
for _ in range(steps):
    step()
torch.cuda.synchronize()
prof = {k: round(v[0] / v[1], 4) for k, v in ws.prof_report().items()}
npx = n * h * w
print(f"{name} [{os.environ.get('VARIANT', '')}]: step median {ms[len(ms) // 2]:.4f} ms min {ms[0]:.4f} -> {npx / ms[len(ms) // 2] / 1e6:.1f} Gpx/s  {prof}", flush=True)
