"""512^3 slab-sharded control-point solve over the ranks of a torchrun job (NCCL all-to-all
between the local sweeps and the axis-0 sweep).  torchrun --nproc-per-node N scripts/sharded_solve_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bsplineinterpolation_b200.distributed import ShardedSolve3D, shard_range
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
shape = (n, n, n)
sh = ShardedSolve3D(3, shape, [(0.0, 1.0)] * 3, device=local)
b, e = shard_range(n, rank, world)
f = torch.rand((e - b, n, n), dtype=torch.float64, device="cuda")
for _ in range(3):
    sh.solve(f)
dist.barrier(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    a = torch.cuda.Event(enable_timing=True); z = torch.cuda.Event(enable_timing=True)
    dist.barrier(); a.record(); sh.solve(f); z.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(z)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(t.item())
if rank == 0:
    print("sharded %d^3 solve on %d GPUs: %.3f ms (median of max over ranks), %s" % (n, world, sorted(ts)[2], ["%.2f" % x for x in ts]))
sh.enable_fused_exchange()
for _ in range(3):
    sh.solve_fused(f)
dist.barrier(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    a = torch.cuda.Event(enable_timing=True); z = torch.cuda.Event(enable_timing=True)
    dist.barrier(); a.record(); sh.solve_fused(f); z.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(z)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(t.item())
if rank == 0:
    print("fused sweep+exchange %d^3 solve on %d GPUs: %.3f ms (median of max over ranks), %s" % (
        n, world, sorted(ts)[2], ["%.2f" % x for x in ts]))
sh.close_fused_exchange()
dist.destroy_process_group()
