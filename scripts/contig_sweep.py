"""Time one in-place sweep along each axis of a 512^3 device array (contiguous axis = 2) and check
that the contiguous-axis kernel agrees bit for bit with the strided kernel run on the transposed array."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B
from scripts.quick_bench import timeit

for per in (False, True):
    for n in (512, 256, 100):
        shape = (n, n, n)
        t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 3, [per] * 3)
        x = torch.rand(shape, dtype=torch.float64, device="cuda")
        a = x.clone()
        t.sweep_axis(2, a, [1, n, n], [0, n * n, n], 1)                 # lines along the contiguous axis
        bt = x.transpose(1, 2).contiguous()                             # [i][k][j]: axis-2 lines now have stride n
        t.sweep_axis(2, bt, [1, n, n], [0, n * n, 1], n)
        same = torch.equal(a, bt.transpose(1, 2))
        ms_c = timeit(lambda: t.sweep_axis(2, a, [1, n, n], [0, n * n, n], 1))[1]
        ms_s = timeit(lambda: t.sweep_axis(0, a, [1, n, n], [0, n, 1], n * n))[1]
        gb = 2 * x.numel() * 8 / 1e6
        print("n=%d periodic=%s: contiguous-axis sweep %.3f ms (%.0f GB/s 1R+1W)  strided sweep %.3f ms  bit-equal=%s" % (
            n, per, ms_c, gb / ms_c, ms_s, same), flush=True)
