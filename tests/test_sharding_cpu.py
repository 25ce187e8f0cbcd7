"""Multi-GPU host logic on CPU: frame-pair sharding across ranks (world_size 2, gloo).  No data-path
collective exists in inference; the only communication is the bench's timing max-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pwclonet_pylidarslam_b200 import sharding


def test_shard_ranges_cover_and_balance():
    for total in (1, 7, 64, 65, 1000):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            for a, b in zip(parts, parts[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 10
        lo, hi = sharding.shard_range(total, rank, world)
        # each rank "processes" its pairs: pose row = pair id (stands in for the forward)
        local = torch.arange(lo, hi, dtype=torch.float32).reshape(-1, 1, 1).expand(-1, 4, 7).contiguous()
        full = sharding.gather_poses(local, total)
        ms = sharding.max_over_ranks(float(rank + 1))
        if rank == 0:
            np.save(out, np.concatenate([full.numpy().reshape(total, -1)[:, 0], [ms]]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_and_timing_reduce(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "r0.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    np.testing.assert_array_equal(got[:10], np.arange(10))
    assert got[10] == 2.0
