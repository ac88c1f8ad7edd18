"""Measures what the box's PCIe link gives: H2D alone, D2H alone, both at once
(pinned memory, two streams) -- the ceiling of the numpy-API (e2e) throughput."""
import time
import torch

n = 256 << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


def both_chunked(k=16):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1):
            d_a[i * c:(i + 1) * c].copy_(h_a[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2):
            h_b[i * c:(i + 1) * c].copy_(d_b[i * c:(i + 1) * c], non_blocking=True)


for name, fn, nbytes in (("h2d", h2d, n), ("d2h", d2h, n), ("both", both, 2 * n), ("both_chunked", both_chunked, 2 * n)):
    t = timeit(fn)
    print(f"{name:14s} {t*1e3:8.2f} ms  {nbytes/t/1e9:7.1f} GB/s")
