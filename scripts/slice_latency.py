"""Development aid: one 512x512 slice through the cluster path and the band kernels -- device time per round trip (events,
L2 flushed before every step, few steps after a synchronise, like bench.py's `slice` leg) and host time per call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import RoundTrip

dev = torch.device("cuda:0")
for path in ("1", "0", "1", "0"):
    os.environ["PEEB_CLUSTER"] = path
    rt = RoundTrip("slice", 0, dev)
    for _ in range(30):
        rt.step()
    torch.cuda.synchronize()
    ms = [rt.timed_steps(5) / 5 for _ in range(6)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        rt.step()
    host_us = (time.perf_counter() - t0) / 200 * 1e6
    torch.cuda.synchronize()
    print(f"PEEB_CLUSTER={path}: device us per round trip (6 x mean of 5): {[round(m * 1e3, 1) for m in ms]}; host us per step (200 back to back): {host_us:.1f}")
