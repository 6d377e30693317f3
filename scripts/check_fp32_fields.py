"""One-off GPU check of the float instantiation of the field-streaming kernel (numpy only, a few seconds)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bsplineinterpolation_b200 as pkg
rng = np.random.default_rng(84)
F, shape, Q = 9, (40, 52), 20000
fields = rng.standard_normal((F,) + shape)
rg = [(0.0, 1.0), (-1.0, 2.0)]
for order, per in [(3, (False, True)), (4, (True, False))]:
    fn32 = pkg.InterpolationFunctionTemplate(order, shape, rg, per, dtype=np.float32).interpolate(fields.astype(np.float32))
    fn64 = pkg.InterpolationFunctionTemplate(order, shape, rg, per).interpolate(fields)
    pts = np.array([0.0, -1.0]) + rng.uniform(0, 1, (Q, 2)) * np.array([1.0, 3.0])
    allv = fn32.evaluate_fields(pts.astype(np.float32))
    ref = fn64.evaluate_fields(pts)
    worst_single = max(np.abs(allv[k] - fn32.evaluate(pts.astype(np.float32), field=k)).max() for k in range(F))
    rel = np.sqrt(((allv - ref) ** 2).sum() / (ref ** 2).sum())
    print("order %d per %s: max |fields - single| = %.3e, rel_err vs f64 = %.3e, max abs vs f64 = %.3e (scale %.2f)"
          % (order, per, worst_single, rel, np.abs(allv - ref).max(), np.abs(ref).max()), flush=True)
