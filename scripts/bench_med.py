"""N1: device-resident timing of the causal (MED) PEE path on the ct512 workload shape, beside the
rhombus path.  usage: python scripts/bench_med.py [n h w bit_depth T]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from codec_tcc_b200 import _cabi, device as D
from codec_tcc_b200.synth import synth_batch

a = [int(x) for x in sys.argv[1:]]
n, h, w, bd, T = (a + [512, 512, 512, 16, 96][len(a):])[:5]
dev = torch.device("cuda:0")
imgs = synth_batch(n, h, w, (1 << bd) - 1, 2)
d_imgs = torch.from_numpy(imgs.view(np.int16) if imgs.dtype == np.uint16 else imgs).to(dev)
stride = D.payload_stride(h * w)
d_pays = torch.from_numpy(np.random.default_rng(7).integers(0, 256, (n, stride), dtype=np.uint8)).to(dev)
out = {}
for pred in ("rhombus", "med"):
    big = np.full(n, h * w, np.int64)
    _, _, d_info = D.pee_embed_device(d_imgs, d_pays, big, T, bd, marked=False, lm=False, predictor=pred)
    cap = d_info[:, 2].cpu().numpy().astype(np.int64)
    d_marked = torch.empty_like(d_imgs); d_lm = torch.empty((n, h, (w + 7) // 8), dtype=torch.uint8, device=dev)
    d_rec = torch.empty_like(d_imgs); d_out = torch.empty((n, stride), dtype=torch.uint8, device=dev)

    def emb():
        D.pee_embed_device(d_imgs, d_pays, cap, T, bd, marked=d_marked, lm=d_lm, predictor=pred)

    def ext():
        D.pee_extract_device(d_marked, d_lm, T, cap, bd, payload_out=d_out, recovered=d_rec, predictor=pred)

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    te, tx = timed(emb), timed(ext)
    assert torch.equal(d_rec, d_imgs)
    npx = n * h * w
    out[pred] = {"embed_ms": te, "extract_ms": tx, "roundtrip_gpixel_s": npx / ((te + tx) * 1e-3) / 1e9,
                 "capacity_bpp": float(cap.sum()) / npx, "algorithmic_gb_s": 8.5 * npx / ((te + tx) * 1e-3) / 1e9}
print(json.dumps({"workload": f"{n} x {h}x{w} {bd}-bit T={T}", **out}, indent=1))
