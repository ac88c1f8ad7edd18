"""Development aid: where the time of peeb_lsb_extract goes -- the whole call (clear + kernel, as scripts/bench_rows.py
times it) against the kernel alone (the workspace's event pair around the launch).  18 Mbit from s = 9 planes of a
3000x3000 16-bit image, segment lengths like the hybrid embedder's."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codec_tcc_b200 import _cabi  # noqa: E402

dev = torch.device("cuda:0")
L = _cabi.lib()
ws = _cabi.workspace(0)
n, s = 3000 * 3000, 9
rng = np.random.default_rng(5)
stego = torch.from_numpy(rng.integers(0, 4096, n, dtype=np.uint16).view(np.int16)).to(dev)
ln = np.array([n] + [n // 8] * 8, dtype=np.int64)
st = np.array([123457] + [0] * 8, dtype=np.int64)
off = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int64)
tot = int(ln.sum())
bits = torch.zeros((tot + 7) // 8 + 16, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream


def call():
    _cabi.check(L.peeb_lsb_extract(ws.handle, stego.data_ptr(), n, 2, s, st.ctypes.data, ln.ctypes.data, off.ctypes.data, tot,
                                   bits.data_ptr(), stream))


for _ in range(3):
    call()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(20_000_000)  # the calls queue up behind ~10 ms of spinning: device time, not the host's enqueue rate
a.record()
for _ in range(50):
    call()
b.record()
torch.cuda.synchronize()
whole = a.elapsed_time(b) / 50
ws.prof_enable(True)
torch.cuda._sleep(20_000_000)
for _ in range(50):
    call()
torch.cuda.synchronize()
rep = {k: v[0] / v[1] for k, v in ws.prof_report().items()}
ws.prof_enable(False)
# reference values for the check: plane bits of the first words
got = bits.cpu().numpy()
px = stego.cpu().numpy().view(np.uint16)
want0 = np.packbits(((np.roll(px, -int(st[0]))[:64] >> 0) & 1).astype(np.uint8))
assert np.array_equal(got[:8], want0), (got[:8], want0)
alg = 2 * tot + tot / 8
print(json.dumps({"call_us": whole * 1e3, "kernel_us": {k: v * 1e3 for k, v in rep.items()}, "algorithmic_MB": alg / 1e6,
                  "gb_s_call": alg / whole / 1e6, "gb_s_kernel": {k: alg / v / 1e6 for k, v in rep.items()}}))
