"""Phase times of the fused sharded 512^3 solve (BSPL_SHARDED_TIMING=1), under torchrun:
   BSPL_SHARDED_TIMING=1 python -m torch.distributed.run --nproc-per-node N scripts/sharded_phases.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bsplineinterpolation_b200 as B
from bsplineinterpolation_b200.distributed import ShardedSolve3D
from bench import smooth_field_slab, SOLVE_MESH
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
if len(sys.argv) > 1:
    B.set_sweep_path(sys.argv[1])
sh = ShardedSolve3D(3, SOLVE_MESH, [(0.0, 1.0)] * 3, device=local)
b, e = sh.slab0[rank], sh.slab0[rank + 1]
f = torch.from_numpy(smooth_field_slab(SOLVE_MESH, b, e)).cuda(local)
for it in range(6):
    dist.barrier(); torch.cuda.synchronize()
    sh.solve_fused(f)
torch.cuda.synchronize()
# pack / finish (the NCCL variant's halves) with events
def tm(fn):
    dist.barrier(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); c = torch.cuda.Event(enable_timing=True)
    a.record(); fn(); c.record(); torch.cuda.synchronize(); return a.elapsed_time(c)
for it in range(3):
    t = tm(lambda: sh.solve(f))
if rank == 0:
    print("nccl variant whole: %.3f ms" % t, file=sys.stderr)
sh.close()
dist.destroy_process_group()
