"""Template construction time of long 1-D axes: uniform (compact host factors), non-uniform on the device, non-uniform
on the host (BSPL_DEVICE_LU_MIN=0), periodic and not.  python scripts/template_build_times.py [log2 n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bsplineinterpolation_b200 as B
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << lg
rng = np.random.default_rng(1)
x = np.arange(n + 1, dtype=np.float64) + rng.uniform(-0.3, 0.3, n + 1); x[0] = 0; x[-1] = n
torch.zeros(1).cuda()
def build(order, coords, per):
    t0 = time.perf_counter()
    t = B.InterpolationFunctionTemplate(order, (n,), [coords], [per])
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0), t.axis_info(0)[2]
for order in (3, 5):
    for per in (False, True):
        xc = x[: n + per]
        os.environ["BSPL_DEVICE_LU_MIN"] = "16384"
        build(order, xc, per)
        d, on_dev = build(order, xc, per)
        os.environ["BSPL_DEVICE_LU_MIN"] = "0"
        h, _ = build(order, xc, per)
        u, _ = build(order, (0.0, 1.0), per)
        print("n 2^%d order %d %s: non-uniform %8.1f ms (%s) | host LU %8.1f ms | uniform axis %6.1f ms"
              % (lg, order, "periodic    " if per else "non-periodic", d, "device" if on_dev else "host", h, u))
