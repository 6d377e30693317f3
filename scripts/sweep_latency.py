"""How long one strided sweep takes as a function of the lines in flight (n = 512, cubic,
non-periodic): few lines = L2-resident working set, the time is the dependent chain of one line;
many lines = HBM-bound.  python scripts/sweep_latency.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B

n = 512
t = B.InterpolationFunctionTemplate(3, (n, n, n), [(0.0, 1.0)] * 3, [False] * 3)


def tm(fn, reps=7):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for lines in (32 * 148, 64 * 148, 128 * 148, 256 * 148, 512 * 148, 1024 * 148, 512 * 512):
    w = torch.rand((n, lines), dtype=torch.float64, device="cuda")
    ms = tm(lambda: t.sweep_axis(0, w, (1, 1, lines), (0, 0, 1), lines))
    mb = n * lines * 8 / 1e6
    print("lines %7d (%4d per SM, %7.1f MB): %.4f ms  -> %.2f us per 148*32 lines, %.0f GB/s (1R+1W)"
          % (lines, lines // 148, mb, ms, ms * 1e3 / (lines / (148 * 32)), 2 * mb / ms))
    del w
