"""Solve times of several shapes under the two sweep routes (device resident, CUDA events, median of 7)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bsplineinterpolation_b200 as B
def tm(fn, reps=7):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
cases = [("cfg5 4096 x 128^2", 3, (128, 128), [False] * 2, 4096), ("256^3", 3, (256,) * 3, [False] * 3, 1),
         ("384^3", 3, (384,) * 3, [False] * 3, 1), ("512^3", 3, (512,) * 3, [False] * 3, 1),
         ("512^3 periodic", 3, (512,) * 3, [True] * 3, 1), ("512^3 quintic", 5, (512,) * 3, [False] * 3, 1),
         ("128^3", 3, (128,) * 3, [False] * 3, 1), ("64 x 2048^2", 3, (64, 2048, 2048), [False] * 3, 1),
         ("8192^2", 3, (8192, 8192), [False] * 2, 1), ("16 x 1024^2 fields", 3, (1024, 1024), [False] * 2, 16),
         ("512^3 float", 3, (512,) * 3, [False] * 3, -1)]
for name, order, shape, per, fields in cases:
    dt = torch.float32 if fields < 0 else torch.float64
    nf = abs(fields)
    row = []
    for path in ("lines", "tiled", "auto"):
        B.set_sweep_path(path)
        t = B.InterpolationFunctionTemplate(order, shape, [(0.0, 1.0)] * len(shape), per,
                                            dtype=("float32" if fields < 0 else "float64"))
        f = torch.rand(((nf,) if nf > 1 else ()) + tuple(shape), dtype=dt, device="cuda")
        fn = t.interpolate(f)
        row.append(tm(lambda: t.interpolate(f, into=fn)))
        del fn, f, t
        torch.cuda.empty_cache()
    print("%-22s lines %.3f ms | tiled %.3f ms | auto %.3f ms" % (name, *row))
B.set_sweep_path("auto")
