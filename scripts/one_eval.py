"""One evaluation call on a 256^3 cubic spline (for ncu): python scripts/one_eval.py <path> <log2 Q> [grad]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bsplineinterpolation_b200 as B
path = sys.argv[1] if len(sys.argv) > 1 else "auto"
Q = 1 << int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 24
n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
shape = (n, n, n)
t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 3)
fn = t.interpolate(torch.rand(shape, dtype=torch.float64, device="cuda"))
pts = torch.rand((Q, 3), dtype=torch.float64, device="cuda")
if len(sys.argv) > 4 and sys.argv[4] == "sorted":
    key = torch.zeros(Q, dtype=torch.int64, device="cuda")
    for d in range(3):
        key = key * n + (pts[:, d] * (n - 1)).long()
    pts = pts[torch.argsort(key)].contiguous()
out = torch.empty((Q, 4), dtype=torch.float64, device="cuda")
B.set_eval_path(path)
for _ in range(3):
    fn.value_grad(pts, out=out)
torch.cuda.synchronize()
