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
for n in (64, 256, 512):
    shape = (n, n, n)
    t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 3)
    fn = t.interpolate(torch.rand(shape, dtype=torch.float64, device="cuda"))
    for lq in range(14, 24):
        Q = 1 << lq
        pts = torch.rand((Q, 3), dtype=torch.float64, device="cuda")
        out = torch.empty((Q, 4), dtype=torch.float64, device="cuda")
        r = {}
        for path in ("direct", "binned"):
            B.set_eval_path(path)
            r[path] = tm(lambda: fn.value_grad(pts, out=out))
        B.set_eval_path("auto")
        print("n=%d Q=2^%d direct %.4f ms binned %.4f ms  -> %s" % (n, lq, r["direct"], r["binned"],
              "binned" if r["binned"] < r["direct"] else "direct"), flush=True)
