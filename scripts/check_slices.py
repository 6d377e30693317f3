import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B
shape = (64, 64, 64)
fn = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 3).interpolate(torch.rand(shape, dtype=torch.float64, device="cuda"))
Q = (1 << 28) + 5000
pts = torch.rand((Q, 3), dtype=torch.float64, device="cuda")
out = fn.evaluate(pts)
torch.cuda.synchronize()
idx = torch.cat([torch.arange(0, 3000, device="cuda"), torch.arange(Q - 6000, Q, device="cuda")])
B.set_eval_path("direct")
ref = fn.evaluate(pts[idx].contiguous())
print("sliced batch max diff vs direct:", float((out[idx] - ref).abs().max()), "launches", B.launch_count())
