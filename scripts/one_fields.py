"""cfg5 many-field evaluation, query-major and field-major (for ncu): python scripts/one_fields.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B
F, shape, Q = 4096, (128, 128), 1 << 20
t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 2)
fn = t.interpolate(torch.rand((F,) + shape, dtype=torch.float64, device="cuda"))
pts = torch.rand((Q, 2), dtype=torch.float64, device="cuda")
out = torch.empty((Q, F), dtype=torch.float64, device="cuda")
for _ in range(2):
    fn.evaluate_fields(pts, out=out, layout="query_major")
torch.cuda.synchronize()
