"""Development aid: device time of the threshold-selection histogram kernel (peeb_pee_hist_batch) on a noise batch and on
tiles of the reference's images/pe.dcm.  usage: python scripts/hist_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from codec_tcc_b200 import _cabi
from codec_tcc_b200.synth import synth_batch

dev = torch.device("cuda:0")
ws = _cabi.workspace(0)
L = _cabi.lib()
pe = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fixtures.npz"))["pe"]
for name, imgs, bd in (("noise16", synth_batch(256, 512, 512, 65535, 2), 16), ("pe12", np.stack([pe] * 256), 12)):
    d = torch.from_numpy(imgs.view(np.int16)).to(dev)
    n, h, w = imgs.shape
    tmax = 1 << (bd - 1)
    hist = torch.zeros((n, 4 * tmax), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    def run():
        _cabi.check(L.peeb_pee_hist_batch(ws.handle, d.data_ptr(), h * w * 2, n, h, w, 2, bd, hist.data_ptr(), st))
    for _ in range(3):
        run()
    ws.prof_enable(True)
    for _ in range(10):
        run()
    torch.cuda.synchronize()
    p = ws.prof_report(); ws.prof_enable(False)
    ms = p["pee_hist"][0] / p["pee_hist"][1]
    print(f"{name} [{os.environ.get('VARIANT', '')}]: hist kernel {ms * 1e3:.1f} us for {n} slices -> {n * h * w / ms / 1e6:.1f} Gpx/s; checksum {int(hist.sum().item())}")
