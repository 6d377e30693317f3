"""One sweep per axis, thread-per-line against the L2-resident tiled kernel, on smooth and on random data."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bsplineinterpolation_b200 as B
from bench import smooth_field_np
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
per = len(sys.argv) > 2 and sys.argv[2] == "periodic"
t = B.InterpolationFunctionTemplate(3, (n, n, n), [(0.0, 1.0)] * 3, [per] * 3)
geoms = {2: ((1, n, n), (0, n * n, n), 1), 1: ((1, n, n), (0, n * n, 1), n), 0: ((1, 1, n * n), (0, 0, 1), n * n)}
for name, f in (("smooth", smooth_field_np((n, n, n))), ("random", np.random.default_rng(1).standard_normal((n, n, n)))):
    for axis in (2, 1, 0):
        out = {}
        for path in ("lines", "tiled"):
            B.set_sweep_path(path)
            w = torch.from_numpy(f).cuda()
            t.sweep_axis(axis, w, *geoms[axis])
            torch.cuda.synchronize()
            out[path] = w.cpu().numpy()
        d = out["lines"] != out["tiled"]
        msg = "%s axis %d: differing %d, max abs diff %.3e" % (name, axis, int(d.sum()), np.abs(out["lines"] - out["tiled"]).max())
        if d.any():
            idx = np.argwhere(d)
            msg += " first %s min %s max %s" % (idx[0].tolist(), idx.min(0).tolist(), idx.max(0).tolist())
            # distribution along the swept axis
            along = np.bincount(idx[:, axis], minlength=n)
            msg += " | rows with differences: %d (first %d last %d)" % ((along > 0).sum(), np.argmax(along > 0), n - 1 - np.argmax(along[::-1] > 0))
        print(msg)
B.set_sweep_path("auto")
