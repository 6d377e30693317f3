"""The bench's smooth 512^3 field through every solve route, against each other and the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bsplineinterpolation_b200 as B
from bench import smooth_field_np, SOLVE_MESH
from oracle.pyoracle import OracleSpline
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
shape = (n, n, n)
f = smooth_field_np(shape)
fd = torch.from_numpy(f).cuda()
res = {}
for path in ("lines", "tiled", "auto"):
    B.set_sweep_path(path)
    fn = B.InterpolationFunction(3, fd, [(0.0, 1.0)] * 3, [False] * 3)
    res[path] = fn.control_points()
    del fn
B.set_sweep_path("auto")
o = OracleSpline(3, shape, [False] * 3, lo=[0, 0, 0], hi=[1, 1, 1], f=f, nthreads=16).control_points()
for k, v in res.items():
    d = v != o
    print(k, "equal to oracle:", not d.any(), "differing:", int(d.sum()), "max abs diff %.3e" % np.abs(v - o).max(),
          "nan:", int(np.isnan(v).sum()))
    if d.any():
        idx = np.argwhere(d)
        print("   first differing indices", idx[:5].tolist(), "axis extents of differing", idx.min(0).tolist(), idx.max(0).tolist())
        i = tuple(idx[0]); print("   values", v[i], o[i])
