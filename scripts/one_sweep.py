"""Three strided sweeps (axis 1) of a 512^3 array, for ncu: python scripts/one_sweep.py [n] [periodic]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
per = len(sys.argv) > 2 and sys.argv[2] == "periodic"
t = B.InterpolationFunctionTemplate(3, (n, n, n), [(0.0, 1.0)] * 3, [per] * 3)
w = torch.rand((n, n, n), dtype=torch.float64, device="cuda")
for _ in range(3):
    t.sweep_axis(1, w, (1, n, n), (0, n * n, 1), n)
torch.cuda.synchronize()
