"""Two 512^3 cubic solves (for ncu: skip the first with -s): python scripts/one_solve.py [n] [periodic]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
per = len(sys.argv) > 2 and sys.argv[2] == "periodic"
shape = (n, n, n)
t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 3, [per] * 3)
f = torch.rand(shape, dtype=torch.float64, device="cuda")
fn = t.interpolate(f)
t.interpolate(f, into=fn)
torch.cuda.synchronize()
