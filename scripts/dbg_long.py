import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsplineinterpolation_b200 as B
from oracle.pyoracle import OracleSpline
case = sys.argv[1]
if case == "q5":
    rng = np.random.default_rng(905)
    n = 200_003
    f = np.cos(np.arange(n) * 0.001) + 0.3 * rng.standard_normal(n)
    o = OracleSpline(5, (n,), [False], lo=[-1.0], hi=[3.0], f=f)
    fn = B.InterpolationFunction(5, f, [(-1.0, 3.0)], [False])
    c, ref = fn.control_points(), o.control_points()
    d = np.abs(c - ref)
    print("max diff", d.max(), "at", d.argmax(), "scale", np.abs(ref).max(), "mismatch frac", (c != ref).mean())
    bad = np.nonzero(d > 1e-13)[0]
    print("bad idx", bad[:10], bad[-10:], len(bad))
else:
    import torch
    n = int(case)
    t = B.InterpolationFunctionTemplate(5, (n,), [(0.0, 1.0)], [True])
    fn = t.interpolate(torch.rand(n, dtype=torch.float64, device="cuda"))
    torch.cuda.synchronize()
    print("done", n, flush=True)
