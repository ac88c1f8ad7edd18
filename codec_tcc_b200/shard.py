"""Image-wise sharding of slice batches / series over the GPUs of one box
(SURVEY.md section 8e).  Images are independent units, so ranks take contiguous
blocks and there is NO collective on the data path; the only exchange is an
optional final gather of the per-image statistics (a few int64 per image)."""
from __future__ import annotations

import os

import numpy as np


def rank_world():
    """(rank, world) from torch.distributed if initialised, else the launcher env."""
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def partition(n_units: int, world: int, rank: int):
    """Contiguous block [lo, hi) of rank ``rank``: sizes differ by at most one,
    earlier ranks take the larger blocks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def partition_grid(n_images: int, n_params: int, world: int, rank: int):
    """The threshold sweep's unit is the (image, T) pair:
    flat index u = image * n_params + param, contiguous blocks of it."""
    lo, hi = partition(n_images * n_params, world, rank)
    u = np.arange(lo, hi)
    return u // max(n_params, 1), u % max(n_params, 1)


def bind_to_gpu_numa(device_index: int) -> list:
    """Pins the calling process to the CPU cores that are local to GPU
    ``device_index`` (NVML's CPU affinity mask), so that pinned host buffers are
    allocated on the GPU's own NUMA node and several ranks of one box do not
    push their PCIe traffic through the socket interconnect.  Returns the cores
    (empty list when NVML or the affinity call is unavailable: then nothing is
    changed)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cores = [64 * k + b for k, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 -- best effort, never fatal
        return []


def gather_stats(local_stats, n_total: int | None = None):
    """all_gather of per-image int64 statistics (rows) -> (n_total, k) on every
    rank.  Works on the NCCL backend (CUDA tensors) and on gloo (CPU)."""
    import torch
    import torch.distributed as dist

    t = local_stats if isinstance(local_stats, torch.Tensor) else torch.as_tensor(np.asarray(local_stats))
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    world = dist.get_world_size()
    if dist.get_backend() == "nccl" and not t.is_cuda:
        t = t.cuda()
    counts = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device))
    counts = [int(c.item()) for c in counts]
    width = t.shape[1] if t.dim() > 1 else 1
    pad = max(counts)
    buf = torch.zeros((pad, width), dtype=t.dtype, device=t.device)
    buf[: t.shape[0]] = t.reshape(t.shape[0], width)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    if n_total is not None and out.shape[0] != n_total:
        raise RuntimeError(f"gathered {out.shape[0]} rows, expected {n_total}")
    return out


def sweep_sharded(imgs, payloads, T_values, bit_depth=None, device=None):
    """Threshold sweep of a whole series over the ranks of one box: the (image, T) grid is split in
    contiguous blocks, every rank embeds its pairs on its own GPU (no data-path collective) and the
    per-pair statistics are gathered at the end.  -> (n_images * n_T, 10) int64 on every rank:
    columns {image, T, n_bits, capacity, cap0, cap1, n_flagged, sse, status, rank}."""
    from . import pee

    rank, world = rank_world()
    Ts = np.asarray(list(T_values), dtype=np.int32)
    n_images = len(imgs)
    img_idx, t_idx = partition_grid(n_images, Ts.size, world, rank)
    info = pee.pee_sweep_pairs(imgs, payloads, img_idx, Ts[t_idx], bit_depth, device=device) if img_idx.size else \
        np.zeros((0, 8), np.int64)
    local = np.concatenate([img_idx[:, None], info, np.full((img_idx.size, 1), rank, np.int64)], axis=1)
    full = gather_stats(local, n_images * Ts.size)
    return full.cpu().numpy() if hasattr(full, "cpu") else np.asarray(full)
