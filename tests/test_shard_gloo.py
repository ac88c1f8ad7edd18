"""CPU: the multi-GPU plumbing (image-wise sharding, final stats gather) with
world_size 2 on the gloo backend -- the data path itself has no collective."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from codec_tcc_b200 import shard


def test_partition_covers_everything_once():
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.partition(n, world, r)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard.partition(n, world, r)[1] - shard.partition(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.partition(4, 2, 2)


def test_partition_grid_is_image_major():
    imgs, params = [], []
    for r in range(3):
        i, p = shard.partition_grid(5, 4, 3, r)
        imgs.extend(i.tolist()); params.extend(p.tolist())
    assert imgs == [u // 4 for u in range(20)] and params == [u % 4 for u in range(20)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert shard.rank_world() == (rank, world)
        lo, hi = shard.partition(n_total, world, rank)
        # stand-in for the per-image info rows a rank's kernels produce
        local = np.stack([np.arange(lo, hi), np.arange(lo, hi) ** 2, np.full(hi - lo, rank)], axis=1).astype(np.int64)
        allstats = shard.gather_stats(local, n_total).numpy()
        assert allstats.shape == (n_total, 3)
        assert np.array_equal(allstats[:, 0], np.arange(n_total))
        assert np.array_equal(allstats[:, 1], np.arange(n_total) ** 2)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), allstats)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 64])
def test_gather_stats_world2_gloo(tmp_path, n_total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b) and a.shape[0] == n_total
