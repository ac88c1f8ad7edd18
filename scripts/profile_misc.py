"""Development aid: one call each of the round-2 kernels that the headline step does not launch, for an ncu capture
(-k regex:'pee2_hist|pee2_pick|pee2_cluster|pbr_|moments_f64'): threshold selection on a 64-slice batch, the cluster path on
one slice, the PBR1 coding of 144 MB of bitmaps, the float64 metrics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from codec_tcc_b200 import device as D, mse
from codec_tcc_b200.synth import synth_batch

dev = torch.device("cuda:0")
n, h, w, bd = 64, 512, 512, 16
imgs = synth_batch(n, h, w, 65535, 21)
d_imgs = torch.from_numpy(imgs.view(np.int16)).to(dev)
stride = D.payload_stride(h * w)
d_pays = torch.from_numpy(np.random.default_rng(3).integers(0, 256, (n, stride), dtype=np.uint8)).to(dev)
nb = np.full(n, 40000, np.int64)
for _ in range(2):
    m, lm, info = D.pee_embed_device(d_imgs, d_pays, nb, None, bd)                      # pee2_hist, pee2_pick_T, rounds
T1 = int(info[0, 0].item())
for _ in range(2):
    m1, lm1, i1 = D.pee_embed_device(d_imgs[:1], d_pays[:1], nb[:1], T1, bd)            # cluster embed
    D.pee_extract_device(m1, lm1, T1, nb[:1], bd)                                       # cluster extract
maps = (torch.rand(4 * 4 * 3000 * 3000, device=dev) < 0.02).to(torch.uint8)
for _ in range(2):
    blob = D.bitmap_encode_device(maps)                                                 # pbr_pack / scan / compact
    D.bitmap_decode_device(blob, maps.numel())                                          # pbr_plan / expand
a = imgs[0].astype(np.float64) + 0.25
mse.AnalisadorMSE().calcular_mse(a, imgs[1])                                            # moments_f64
torch.cuda.synchronize()
print("ok", T1, blob.numel())
