import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bsplineinterpolation_b200 as B
Q = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
shape = (256, 256, 256)
t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 3)
fn = t.interpolate(torch.rand(shape, dtype=torch.float64, device="cuda"))
hp = torch.empty((Q, 3), dtype=torch.float64, pin_memory=True); hp.uniform_()
ho = torch.empty((Q, 4), dtype=torch.float64, pin_memory=True)
npts, nout = hp.numpy(), ho.numpy()
for path in ("direct", "binned", "auto"):
    B.set_eval_path(path)
    fn.value_grad(npts, out=nout)
    for _ in range(2):
        t0 = time.perf_counter(); fn.value_grad(npts, out=nout); dt = time.perf_counter() - t0
        print("%s pinned: %.1f ms total (%.2f Gpts/s, %.1f GB/s PCIe both ways), kernels %.1f ms" % (
            path, dt * 1e3, Q / dt / 1e9, Q * 56 / dt / 1e9, B.last_kernel_ms()), flush=True)
pg = np.random.rand(1 << 24, 3); po = np.empty((1 << 24, 4))
fn.value_grad(pg, out=po)
t0 = time.perf_counter(); fn.value_grad(pg, out=po); dt = time.perf_counter() - t0
print("pageable 2^24: %.1f ms (%.2f Gpts/s)" % (dt * 1e3, (1 << 24) / dt / 1e9))
