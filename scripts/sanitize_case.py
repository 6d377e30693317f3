"""Small end-to-end case for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bsplineinterpolation_b200 as B
rng = np.random.default_rng(0)
for order, per in ((3, [False, True, False]), (2, [True, False, False]), (5, [False, False, True])):
    shape = (40, 37, 45)
    f = rng.standard_normal((2,) + shape)
    t = B.InterpolationFunctionTemplate(order, shape, [(0.0, 1.0)] * 3, per)
    fn = t.interpolate(torch.from_numpy(f).cuda())
    pts = torch.from_numpy(rng.uniform(-0.2, 1.2, (30000, 3))).cuda()
    for path in ("direct", "binned"):
        B.set_eval_path(path)
        v = fn.value_grad(pts); e = fn.evaluate(pts, derivatives=[1, 0, 1], field=1)
    B.set_eval_path("auto")
    plan = fn.eval_proxy(pts); plan(fn, value_grad=True, device_out=True)
    fn.evaluate_fields(pts)
# long 1-D (chunked, cyclic), 2-D transposed-first path, contiguous small-batch kernel
B.InterpolationFunction(5, rng.standard_normal(60000), [(0.0, 1.0)], [True])
B.InterpolationFunction(3, rng.standard_normal((300, 200)), [(0.0, 1.0)] * 2, [False, True])
B.InterpolationFunction(3, rng.standard_normal((40, 50)), [(0.0, 1.0)] * 2, [True, False])
torch.cuda.synchronize()
print("sanitize case done")
