"""Per-sweep kernel times of a 3-D solve via CUDA events around each interpolate (ncu-free)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
per = [len(sys.argv) > 2 and sys.argv[2] == "periodic"] * 3
t = B.InterpolationFunctionTemplate(3, (n, n, n), [(0.0, 1.0)] * 3, per)
w = torch.rand((n, n, n), dtype=torch.float64, device="cuda")
def tm(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
print("axis 2 (contiguous) %.3f ms" % tm(lambda: t.sweep_axis(2, w, (1, n, n), (0, n * n, n), 1)))
print("axis 1 (stride n)   %.3f ms" % tm(lambda: t.sweep_axis(1, w, (1, n, n), (0, n * n, 1), n)))
print("axis 0 (stride n^2) %.3f ms" % tm(lambda: t.sweep_axis(0, w, (1, 1, n * n), (0, 0, 1), n * n)))
f = torch.rand((n, n, n), dtype=torch.float64, device="cuda")
fn = t.interpolate(f)
print("interpolate         %.3f ms" % tm(lambda: t.interpolate(f, into=fn)))
c = torch.empty_like(f)
print("device copy (1R+1W) %.3f ms" % tm(lambda: c.copy_(f)))
