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
# TMA-tiled contiguous sweep: fused with the mesh copy (ragged tiles: 70 lines per run, 72 columns),
# with a shifted slow axis, in float, and in place after a rotating copy is not possible (odd strides fall back)
B.InterpolationFunction(3, rng.standard_normal((66, 70, 72)), [(0.0, 1.0)] * 3, [False, False, False])
B.InterpolationFunction(3, rng.standard_normal((66, 70, 72)), [(0.0, 1.0)] * 3, [True, False, False])
B.InterpolationFunction(5, rng.standard_normal((66, 70, 72)).astype(np.float32), [(0.0, 1.0)] * 3, dtype=np.float32)
B.InterpolationFunction(3, rng.standard_normal((66, 70, 73)), [(0.0, 1.0)] * 3, [False, False, False])
# periodic axes through the fused sweep: wrapped rows patched in, line rotation by the delay registers
B.InterpolationFunction(3, rng.standard_normal((66, 70, 72)), [(0.0, 1.0)] * 3, [True, True, True])
B.InterpolationFunction(4, rng.standard_normal((66, 70, 72)), [(0.0, 1.0)] * 3, [False, True, True])
B.InterpolationFunction(5, rng.standard_normal((20, 260, 48)), [(0.0, 1.0)] * 3, [True, False, True])
# compact factors (axis of 20 000 points) with the chunk-parallel and the strided sweeps
B.InterpolationFunction(3, rng.standard_normal(20000), [(0.0, 1.0)], [True])
B.InterpolationFunction(3, rng.standard_normal((20000, 40)), [(0.0, 1.0)] * 2, [False, True])
tt = B.InterpolationFunctionTemplate(4, (80, 64, 96), [(0.0, 1.0)] * 3, [True, True, True])
xx = torch.rand((80, 64, 96), dtype=torch.float64, device="cuda")
tt.sweep_axis(2, xx, [1, 80, 64], [0, 64 * 96, 96], 1)
# round 2: the L2-resident tiled sweeps forced on small meshes (strided rows, swizzled contiguous boxes, periodic
# strips and tail tiles, zeros that send blocks to the full division), the single-rank sharded plan (tiled exchange
# sweep) and the many-field contraction with query-major results
B.set_sweep_path("tiled")
g = rng.standard_normal((136, 132, 144)); g[5:9] = 0.0
B.InterpolationFunction(3, g, [(0.0, 1.0)] * 3, [False, False, False])
B.InterpolationFunction(3, g, [(0.0, 1.0)] * 3, [True, True, False])
B.InterpolationFunction(5, rng.standard_normal((264, 136)), [(0.0, 1.0)] * 2, [True, True])
B.InterpolationFunction(2, g.astype(np.float32), [(0.0, 1.0)] * 3, dtype=np.float32)
from bsplineinterpolation_b200.distributed import ShardedSolve3D
sh = ShardedSolve3D(3, (40, 136, 64), [(0.0, 1.0)] * 3, [True, False, False])
sh.solve_fused(torch.from_numpy(rng.standard_normal((40, 136, 64))).cuda())
sh.solve(torch.from_numpy(rng.standard_normal((40, 136, 64))).cuda())
sh.close()
B.set_sweep_path("auto")
t5 = B.InterpolationFunctionTemplate(3, (32, 40), [(0.0, 1.0)] * 2)
fn5 = t5.interpolate(torch.from_numpy(rng.standard_normal((64, 32, 40))).cuda())
p5 = torch.from_numpy(rng.uniform(0, 1, (5000, 2))).cuda()
for path in ("gather", "contract"):
    B.set_fields_path(path)
    fn5.evaluate_fields(p5); fn5.evaluate_fields(p5, layout="query_major"); fn5.evaluate_fields(p5, layout="query_major", derivatives=[1, 0])
B.set_fields_path("auto")
torch.cuda.synchronize()
print("sanitize case done")
