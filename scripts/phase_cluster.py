"""Development aid: per-phase time of the cluster embed kernel (small-image path), thread 0 of every CTA.
Needs a library built with -DPEEB_PHASE_TIMING (scripts/build_variant.sh phase -DPEEB_PHASE_TIMING, copied over the
library).  usage: PEEB_CLUSTER=1 python scripts/phase_cluster.py [n h w bit_depth T]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from codec_tcc_b200 import _cabi, device as D
from codec_tcc_b200.synth import synth_batch

NAMES = ["set-up", "band copy wait", "count 0", "exchange 0", "apply 0", "count 1", "exchange 1", "apply 1", "statistics + summary",
         "store drain"]
a = [int(x) for x in sys.argv[1:]]
n, h, w, bd, T = (a + [1, 512, 512, 16, 96][len(a):])[:5]
_cabi.lib()
fn = C.CDLL(_cabi.library_path()).peeb_debug_phases
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
dev = torch.device("cuda:0")
imgs = synth_batch(n, h, w, (1 << bd) - 1, 2)
d_imgs = torch.from_numpy(imgs.view(np.int16) if imgs.dtype == np.uint16 else imgs).to(dev)
stride = D.payload_stride(h * w)
d_pays = torch.from_numpy(np.random.default_rng(7).integers(0, 256, (n, stride), dtype=np.uint8)).to(dev)
_, _, d_info = D.pee_embed_device(d_imgs, d_pays, np.full(n, h * w, np.int64), T, bd, marked=False, lm=False)
cap = d_info[:, 2].cpu().numpy().astype(np.int64)
d_marked = torch.empty_like(d_imgs)
d_lm = torch.empty((n, h, (w + 7) // 8), dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    D.pee_embed_device(d_imgs, d_pays, cap, T, bd, marked=d_marked, lm=d_lm)
buf = (C.c_ulonglong * 32)()
fn(buf, 1)
reps = 10
for _ in range(reps):
    flush.fill_(1)
    D.pee_embed_device(d_imgs, d_pays, cap, T, bd, marked=d_marked, lm=d_lm)
fn(buf, 0)
ctas = 16 * n
print(f"cluster embed phases, {n}x{h}x{w}: average per CTA (of {ctas}) and launch, microseconds at 1.965 GHz")
tot = 0.0
for i, nm in enumerate(NAMES):
    us = buf[i] / reps / ctas / 1965.0
    tot += us
    print(f"  {nm:24s} {us:7.2f}")
print(f"  {'sum':24s} {tot:7.2f}")
