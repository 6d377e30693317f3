"""GPU tests of the staged (per-axis) solve used by the multi-GPU path."""
import os
import socket

import numpy as np
import pytest

from cases import smooth_field
from oracle.pyoracle import OracleSpline

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(34, 29, 41), (34, 30, 48), (40, 136, 64)])
@pytest.mark.parametrize("order,periodic", [(3, (False, False, False)), (3, (True, False, True)),
                                            (4, (True, True, True)), (2, (False, True, False))])
def test_staged_solve_single_rank(lib_built, order, periodic, shape):
    """The sharded plan with one rank (bspl_sharded_solve_*: fused first sweep on TMA tiles when the
    contiguous extent allows it, exchange sweep into its own buffer, last sweep) reproduces
    interpolate() bit for bit, through both exchanges.  The last shape is long enough along axis 1 for
    the tiled exchange sweep (sweep_rows_tma_kernel<EXCH>: bulk tensor stores into the owner's buffer)."""
    import torch
    from bsplineinterpolation_b200.distributed import ShardedSolve3D, shard_range
    rng = np.random.default_rng(31 + order)
    f = smooth_field(shape, rng)
    ranges = [(0.0, 1.0), (-1.0, 2.0), (0.5, 4.0)]
    sh = ShardedSolve3D(order, shape, ranges, periodic)
    assert sh.slab0 == [0, shape[0]] and sh.slab1 == [0, shape[1]]
    ctrl = sh.solve(torch.from_numpy(f).cuda())
    o = OracleSpline(order, shape, periodic, lo=[r[0] for r in ranges], hi=[r[1] for r in ranges], f=f)
    assert np.array_equal(ctrl.cpu().numpy(), o.control_points())
    fn = sh.gather_function(ctrl)
    sh.enable_fused_exchange()
    assert np.array_equal(sh.solve_fused(torch.from_numpy(f).cuda()).cpu().numpy(), o.control_points())
    assert not sh.timed_out()
    sh.close_fused_exchange()
    pts = np.array([r[0] for r in ranges]) + rng.uniform(0, 1, (2000, 3)) * np.array([r[1] - r[0] for r in ranges])
    ref = o.eval(pts)
    assert np.abs(fn(pts) - ref).max() <= 1e-12 * np.abs(ref).max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from bsplineinterpolation_b200.distributed import ShardedSolve3D, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(5)
        ok = True
        # the second shape is long enough along axis 1 for the tiled exchange sweep; 160 rows over two ranks
        # put the ownership boundary (row 80) inside a 32-row tile, which is then stored to both owners
        for shape, periodic in (((50, 37, 44), (True, False, True)), ((66, 160, 96), (False, False, False)),
                                ((66, 160, 96), (True, True, False))):
            f = smooth_field(shape, rng)
            ranges = [(0.0, 1.0)] * 3
            sh = ShardedSolve3D(3, shape, ranges, periodic, device=rank)
            b, e = shard_range(shape[0], rank, world)
            assert [sh.slab0[rank], sh.slab0[rank + 1]] == [b, e]
            ctrl = sh.solve(torch.from_numpy(f[b:e]).cuda(rank))
            fn = sh.gather_function(ctrl)
            o = OracleSpline(3, shape, periodic, lo=[0, 0, 0], hi=[1, 1, 1], f=f)
            ok = ok and np.array_equal(fn.control_points(), o.control_points())
            back = sh.solve(torch.from_numpy(f[b:e]).cuda(rank), back_to_axis0=True)
            ok = ok and np.array_equal(back.cpu().numpy(), o.control_points()[b:e])
            # fused sweep + exchange over peer memory
            sh.enable_fused_exchange()
            b1, e1 = shard_range(shape[1], rank, world)
            for _ in range(2):
                fused = sh.solve_fused(torch.from_numpy(f[b:e]).cuda(rank))
                ok = ok and np.array_equal(fused.cpu().numpy(), o.control_points()[:, b1:e1, :])
            ok = ok and not sh.timed_out()
            sh.close()
        with open(os.path.join(out_dir, "r%d" % rank), "w") as fh:
            fh.write("%d" % ok)
    finally:
        dist.destroy_process_group()


def test_sharded_solve_two_gpus(lib_built, tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert all(open(tmp_path / ("r%d" % r)).read() == "1" for r in range(2))
