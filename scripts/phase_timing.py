"""Development aid: per-phase time of the PEE embed kernel (thread 0 of every CTA, clock64 deltas).
Needs a library built with PEEB_NVCC_EXTRA=-DPEEB_PHASE_TIMING (python -m codec_tcc_b200.build --force).
usage: python scripts/phase_timing.py [n_images h w bit_depth T]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from codec_tcc_b200 import _cabi, device as D
from codec_tcc_b200.synth import synth_batch

NAMES = ["set-up + pass-0 order/bits", "band copy wait", "apply0", "barrier", "count1", "barrier", "total + look-back",
         "apply1", "barrier + store issue + stats", "location map + store drain"]


def main():
    a = [int(x) for x in sys.argv[1:] if not x.startswith('--')]
    n, h, w, bd, T = (a + [512, 512, 512, 16, 96][len(a):])[:5]
    _cabi.lib()
    fn = C.CDLL(_cabi.library_path()).peeb_debug_phases
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    dev = torch.device("cuda:0")
    maxval = (1 << bd) - 1
    imgs = synth_batch(n, h, w, maxval, 2)
    d_imgs = torch.from_numpy(imgs.view(np.int16) if imgs.dtype == np.uint16 else imgs).to(dev)
    stride = D.payload_stride(h * w)
    d_pays = torch.from_numpy(np.random.default_rng(7).integers(0, 256, (n, stride), dtype=np.uint8)).to(dev)
    big = np.full(n, h * w, np.int64)
    _, _, d_info = D.pee_embed_device(d_imgs, d_pays, big, T, bd, marked=False, lm=False)
    cap = d_info[:, 2].cpu().numpy().astype(np.int64)
    d_marked = torch.empty_like(d_imgs)
    d_lm = torch.empty((n, h, (w + 7) // 8), dtype=torch.uint8, device=dev)
    for _ in range(3):
        D.pee_embed_device(d_imgs, d_pays, cap, T, bd, marked=d_marked, lm=d_lm)
    buf = (C.c_ulonglong * 32)()
    fn(buf, 1)
    reps = 5
    for _ in range(reps):
        D.pee_embed_device(d_imgs, d_pays, cap, T, bd, marked=d_marked, lm=d_lm)
    fn(buf, 0)
    if "--extract" in sys.argv:
        d_rec = torch.empty_like(d_imgs); d_out = torch.empty((n, stride), dtype=torch.uint8, device=dev)
        D.pee_extract_device(d_marked, d_lm, T, cap, bd, payload_out=d_out, recovered=d_rec)
        fn(buf, 1)
        for _ in range(reps):
            D.pee_extract_device(d_marked, d_lm, T, cap, bd, payload_out=d_out, recovered=d_rec)
        fn(buf, 0)
        names = ["load", "sweep colour 1", "barrier", "sweep colour 0", "barrier", "assemble+stage", "store"]
        names += [""] * 17 + ["  location map + copies issued", "  tables cleared", "  mbarrier wait", "  stores issued", "  assembled", "  barrier"]
        tot = sum(buf[i] for i in range(32))
        print("extract kernel phases")
        for i, nm in enumerate(names):
            if not nm:
                continue
            print(f"  {nm:22s} {buf[i] / reps:14.0f} total  {100.0 * buf[i] / max(tot, 1):5.1f}%")
        return
    names = NAMES + [""] * 6 + ["  copies issued", "  band counts", "  pass-0 order + bits (first item)", "  mbarrier wait", "  barrier after pass 1", "  stores issued"]
    tot = sum(buf[i] for i in range(32))
    print(f"embed kernel phases, {n}x{h}x{w} bd={bd} T={T}: cycles of thread 0 summed over the CTAs, share")
    for i, nm in enumerate(names):
        if not nm:
            continue
        print(f"  {nm:30s} {buf[i] / reps:14.0f} total  {100.0 * buf[i] / max(tot, 1):5.1f}%")


if __name__ == "__main__":
    main()
