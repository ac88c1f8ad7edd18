"""Development aid: device time of the 65 536-bin histogram (row a5) for 12-bit, smooth 16-bit and uniform
random 16-bit data.  usage: python scripts/hist_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from codec_tcc_b200 import _cabi
from codec_tcc_b200.synth import synth_image

dev = torch.device("cuda:0")
ws = _cabi.workspace(0)
L = _cabi.lib()
st = torch.cuda.current_stream().cuda_stream
h = w = 3000
n_img = 8
d_hist = torch.empty(65536, dtype=torch.int32, device=dev)
d_ones = torch.empty(16, dtype=torch.int64, device=dev)
for label, mk in (("12-bit smooth", lambda k: synth_image(h, w, 4095, 50 + k)),
                  ("16-bit smooth", lambda k: synth_image(h, w, 65535, 50 + k)),
                  ("16-bit uniform random", lambda k: np.random.default_rng(k).integers(0, 65536, (h, w), dtype=np.uint16))):
    imgs = np.stack([mk(k) for k in range(n_img)])
    d = torch.from_numpy(imgs.view(np.int16)).to(dev).reshape(-1)

    def run():
        _cabi.check(L.peeb_hist_planes(ws.handle, d.data_ptr(), d.numel(), 2, d_hist.data_ptr(), d_ones.data_ptr(), st))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ref = np.bincount(imgs.reshape(-1), minlength=65536)
    assert np.array_equal(d_hist.cpu().numpy().astype(np.int64), ref)
    print(f"{label:24s} {ms:.4f} ms  {d.numel() / ms / 1e6:.1f} Gpx/s  {2 * d.numel() / ms / 1e6:.0f} GB/s", flush=True)
