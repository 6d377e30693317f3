"""Ad-hoc device timings used during development (not the contract bench)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bsplineinterpolation_b200 as B

def field(shape):
    g = torch.meshgrid(*[torch.arange(n, dtype=torch.float64, device="cuda") for n in shape], indexing="ij")
    f = torch.ones(shape, dtype=torch.float64, device="cuda")
    for a, n in zip(g, shape):
        f = f * torch.cos(2 * np.pi * a / n - np.pi)
    return f.contiguous()

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts)//2]

def eval_case(shape, order, Q, per=None, sort=False):
    dim = len(shape)
    t = B.InterpolationFunctionTemplate(order, shape, [(0.0, 1.0)] * dim, per)
    fn = t.interpolate(field(shape))
    pts = torch.rand((Q, dim), dtype=torch.float64, device="cuda")
    if sort:
        key = torch.zeros(Q, dtype=torch.int64, device="cuda")
        for d in range(dim):
            key = key * shape[d] + (pts[:, d] * (shape[d]-1)).long()
        pts = pts[torch.argsort(key)].contiguous()
    out = torch.empty((Q, dim + 1), dtype=torch.float64, device="cuda")
    outv = torch.empty((Q,), dtype=torch.float64, device="cuda")
    tv = timeit(lambda: fn.evaluate(pts, out=outv))
    tg = timeit(lambda: fn.value_grad(pts, out=out))
    print("eval %s order %d Q=%d sort=%d: value %.3f ms (%.2f Gpts/s)  value+grad %.3f ms (%.2f Gpts/s)" % (
        shape, order, Q, sort, tv[1], Q / tv[1] / 1e6, tg[1], Q / tg[1] / 1e6), flush=True)

def solve_case(shape, order, per=None, fields=1):
    dim = len(shape)
    t0 = time.time()
    t = B.InterpolationFunctionTemplate(order, shape, [(0.0, 1.0)] * dim, per)
    t1 = time.time()
    f = field(shape)
    if fields > 1: f = f.unsqueeze(0).repeat(fields, *([1] * dim)).contiguous()
    fn = t.interpolate(f)
    ts = timeit(lambda: t.interpolate(f, into=fn), reps=5, warm=1)
    nbytes = f.numel() * 8 * 2 * dim
    print("solve %s order %d per=%s fields=%d: template %.1f ms host; interpolate %.3f ms (min %.3f) -> %.1f GB/s algorithmic" % (
        shape, order, per, fields, (t1 - t0) * 1e3, ts[1], ts[0], nbytes / ts[1] / 1e6), flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "binned":
        for path in ("direct", "binned"):
            B.set_eval_path(path); print("path", path)
            eval_case((256, 256, 256), 3, 1 << 22)
            eval_case((256, 256, 256), 3, 1 << 24)
            eval_case((256, 256, 256), 3, 1 << 26)
            eval_case((256, 256, 256), 3, 1 << 28)
            eval_case((64, 64, 64), 3, 1 << 24)
            eval_case((512, 512, 512), 3, 1 << 26)
            eval_case((256, 256, 256), 5, 1 << 24)
            eval_case((256, 256, 256), 1, 1 << 24)
        B.set_eval_path("auto")
    if which in ("all", "eval"):
        eval_case((256, 256, 256), 3, 1 << 24)
        eval_case((256, 256, 256), 3, 1 << 26)
        eval_case((256, 256, 256), 3, 1 << 24, sort=True)
        eval_case((64, 64, 64), 3, 1 << 24)
        eval_case((1024, 1024), 3, 1 << 24)
        eval_case((1 << 24,), 5, 1 << 24, per=[True])
    if which == "plan":
        shape = (256, 256, 256)
        t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 3)
        fn = t.interpolate(field(shape))
        for lq in (24, 26, 28):
            Q = 1 << lq
            pts = torch.rand((Q, 3), dtype=torch.float64, device="cuda")
            out = torch.empty((Q, 4), dtype=torch.float64, device="cuda")
            plan = fn.eval_proxy(pts)
            ms = timeit(lambda: plan(fn, value_grad=True, out=out))[1]
            ms0 = timeit(lambda: fn.value_grad(pts, out=out))[1]
            print("Q=2^%d: planned value+grad %.3f ms (%.2f Gpts/s); one-shot %.3f ms (%.2f Gpts/s)" % (
                lq, ms, Q / ms / 1e6, ms0, Q / ms0 / 1e6), flush=True)
            del plan, pts, out
    if which == "fields":
        F, shape, Q = 4096, (128, 128), 1 << 20
        t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 2)
        f = torch.rand((F,) + shape, dtype=torch.float64, device="cuda")
        fn = t.interpolate(f)
        print("solve 4096 fields: %.3f ms" % timeit(lambda: t.interpolate(f, into=fn))[1])
        pts = torch.rand((Q, 2), dtype=torch.float64, device="cuda")
        out = torch.empty((F, Q), dtype=torch.float64, device="cuda")
        ms = timeit(lambda: fn.evaluate_fields(pts, out=out), reps=3, warm=1)[1]
        print("evaluate_fields 4096 x 2^20: %.2f ms -> %.2f G(query,field)/s, %.0f GB/s algorithmic (152 B each)" % (
            ms, F * Q / ms / 1e6, F * Q * 152 / ms / 1e6))
    if which == "long":
        solve_case((1 << 24,), 5, per=[True])
        solve_case((1 << 24,), 3)
        solve_case((1 << 20,), 3, per=[True])
        solve_case((1024, 1024), 3, per=[True, True])
        solve_case((4096, 4096), 3, per=[True, False])
    if which in ("all", "solve"):
        solve_case((256, 256, 256), 3)
        solve_case((512, 512, 512), 3)
        solve_case((512, 512, 512), 3, per=[True, True, True])
        solve_case((128, 128), 3, fields=4096)
        solve_case((1024, 1024), 3, per=[True, True])
