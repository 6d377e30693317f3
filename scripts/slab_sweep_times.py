"""Sweep times on one rank's slab of a sharded 512^3 solve (64 planes at 8 GPUs, 256 at 2): thread-per-line
kernels against the tiled ones.  python scripts/slab_sweep_times.py [planes]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B
n = 512
pl = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t = B.InterpolationFunctionTemplate(3, (n, n, n), [(0.0, 1.0)] * 3, [False] * 3)
def tm(fn, reps=9):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
for path in ("lines", "tiled"):
    B.set_sweep_path(path)
    w = torch.rand((pl, n, n), dtype=torch.float64, device="cuda")      # [n0_loc][n1][n2]
    v = torch.rand((n, pl, n), dtype=torch.float64, device="cuda")      # [n0][n1_loc][n2]
    print("%-5s planes %3d: axis 2 (contiguous) %.3f ms | axis 1 (stride n) %.3f ms | axis 0 of [n0][n1_loc][n2] %.3f ms"
          % (path, pl, tm(lambda: t.sweep_axis(2, w, (1, pl, n), (0, n * n, n), 1)),
             tm(lambda: t.sweep_axis(1, w, (1, pl, n), (0, n * n, 1), n)),
             tm(lambda: t.sweep_axis(0, v, (1, 1, pl * n), (0, 0, 1), pl * n))))
B.set_sweep_path("auto")
