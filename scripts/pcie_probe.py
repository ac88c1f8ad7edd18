"""The host<->device copy floor of one box with N ranks copying at once: per rank H2D alone, D2H alone and both
directions at once (pinned memory, one stream per direction, 32 MB chunks like the host wrappers of libpeeb200),
every rank at the same time (barrier before each timed block).  Launched like bench.py:

    python scripts/pcie_probe.py                                  # one rank
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_probe.py

Rank 0 prints ONE JSON line {"world": N, "per_rank_gbs_*": [...], "aggregate_gbs_*": ...}; the e2e figure of bench.py
is compared with "both directions" (profiles/pcie_floor_r02.json collects the lines per N)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

n, chunk = 512 << 20, 32 << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def copies(do_h2d, do_d2h):
    for i in range(0, n, chunk):
        if do_h2d:
            with torch.cuda.stream(s1):
                d_a[i:i + chunk].copy_(h_a[i:i + chunk], non_blocking=True)
        if do_d2h:
            with torch.cuda.stream(s2):
                h_b[i:i + chunk].copy_(d_b[i:i + chunk], non_blocking=True)


def timed(do_h2d, do_d2h, reps=4):
    copies(do_h2d, do_d2h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        copies(do_h2d, do_d2h)
    torch.cuda.synchronize(dev)
    sec = (time.perf_counter() - t0) / reps
    barrier()
    return n * (int(do_h2d) + int(do_d2h)) / sec / 1e9


res = {}
for key, a, b in (("h2d", True, False), ("d2h", False, True), ("both_directions", True, True)):
    g = torch.tensor([timed(a, b)], dtype=torch.float64, device=dev)
    every = [torch.zeros_like(g) for _ in range(world)]
    if world > 1:
        dist.all_gather(every, g)
    else:
        every = [g]
    res[key] = [round(float(x.item()), 2) for x in every]
if rank == 0:
    line = {"world": world, "bytes_per_direction_per_rank": n, "chunk_bytes": chunk}
    for key, v in res.items():
        line["per_rank_gbs_" + key] = v
        line["aggregate_gbs_" + key] = round(sum(v), 1)
    # what one rank gets when all copy in both directions at once: GB/s per direction
    line["per_rank_gbs_both_directions_per_direction"] = round(min(res["both_directions"]) / 2, 2)
    print(json.dumps(line), flush=True)
if world > 1:
    dist.destroy_process_group()
