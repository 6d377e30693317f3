"""CPU-only, world_size 2 over gloo: the host-side partitioning logic of the multi-GPU paths
(query / field sharding and the slab re-shard between axis sweeps).  No kernels run here."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bsplineinterpolation_b200.distributed import (reshard_axis0_to_axis1, reshard_axis1_to_axis0, shard_range,
                                                   shard_sizes)


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 256, 1 << 28, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(total, world)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n0, n1, n2 = shape
        full = torch.arange(n0 * n1 * n2, dtype=torch.float64).view(n0, n1, n2)
        b0, e0 = shard_range(n0, rank, world)
        b1, e1 = shard_range(n1, rank, world)
        y = reshard_axis0_to_axis1(full[b0:e0].contiguous(), n0)
        ok1 = torch.equal(y, full[:, b1:e1, :])
        back = reshard_axis1_to_axis0(y, n1)
        ok2 = torch.equal(back, full[b0:e0])
        # query sharding: every rank evaluates its slice, the union covers the batch once
        q = 1001
        mine = torch.zeros(q)
        qb, qe = shard_range(q, rank, world)
        mine[qb:qe] = 1
        dist.all_reduce(mine)
        ok3 = bool((mine == 1).all())
        with open(os.path.join(result_dir, "r%d" % rank), "w") as fh:
            fh.write("%d %d %d" % (ok1, ok2, ok3))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape", [(8, 6, 5), (7, 9, 4), (3, 2, 11)])
def test_slab_reshard_world2_gloo(tmp_path, shape):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), shape, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / ("r%d" % r)).read() == "1 1 1"
