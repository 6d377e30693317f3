import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B
shape = (256, 256, 256)
t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 3)
fn = t.interpolate(torch.rand(shape, dtype=torch.float64, device="cuda"))
for lq in (26, 28):
    Q = 1 << lq
    pts = torch.rand((Q, 3), dtype=torch.float64, device="cuda")
    out = torch.empty((Q, 4), dtype=torch.float64, device="cuda")
    for _ in range(3): fn.value_grad(pts, out=out)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn.value_grad(pts, out=out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[2]
    print(os.environ.get("BSPL_B200_LIB", "default"), "Q=2^%d: %.3f ms  %.2f Gpts/s" % (lq, ms, Q / ms / 1e6), flush=True)
    del pts, out
