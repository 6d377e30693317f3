#!/usr/bin/env python3
"""Generate tests/golden/ref_outputs_f32.npz: outputs of the UNMODIFIED reference instantiated with
T = U = float (oracle/ref_shim_f32.cpp, plain layout), on seeded inputs.  Build container only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
from cases import smooth_field  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
so = os.path.join(ROOT, "oracle", "_ref", "libintp_ref_f32.so")
subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-O3", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden",
                       "-fvisibility-inlines-hidden", "-fno-gnu-unique", "-Wl,-Bsymbolic",
                       "-DINTP_PERIODIC_NO_DUMMY_POINT", "-I" + REF + "/src/include", "-o", so,
                       os.path.join(ROOT, "oracle", "ref_shim_f32.cpp")])
L = C.CDLL(so)
L.intp_ref32_create.restype = C.c_void_p
fp = C.POINTER(C.c_float)
cases = [(1, 3, (True,), (41,)), (1, 5, (False,), (53,)), (2, 3, (False, True), (23, 19)), (2, 2, (True, False), (17, 29)),
         (3, 3, (False, False, False), (13, 11, 15)), (3, 3, (True, False, True), (12, 14, 10)), (3, 1, (False, True, False), (9, 8, 11))]
out = {}
for c, (dim, order, per, shape) in enumerate(cases):
    rng = np.random.default_rng(700 + c)
    f = smooth_field(shape, rng).astype(np.float32)
    lo = np.zeros(dim, dtype=np.float32); hi = np.arange(1, dim + 1, dtype=np.float32)
    n = np.asarray(shape, dtype=np.uint64); p = np.asarray(per, dtype=np.int32)
    h = L.intp_ref32_create(dim, order, n.ctypes.data_as(C.POINTER(C.c_uint64)), p.ctypes.data_as(C.POINTER(C.c_int)),
                            lo.ctypes.data_as(fp), hi.ctypes.data_as(fp), np.ascontiguousarray(f).ctypes.data_as(fp))
    assert h, (dim, order)
    h = C.c_void_p(h)
    pts = (lo + rng.uniform(-0.15, 1.15, (400, dim)).astype(np.float32) * (hi - lo)).astype(np.float32)
    if not all(per):  # keep non-periodic coordinates inside the range
        for d in range(dim):
            if not per[d]:
                pts[:, d] = np.clip(pts[:, d], lo[d], hi[d])
    vals = np.empty(len(pts), dtype=np.float32); dv = np.empty(len(pts), dtype=np.float32)
    spans = np.empty(pts.shape, dtype=np.int64)
    L.intp_ref32_eval(h, pts.ctypes.data_as(fp), C.c_uint64(len(pts)), vals.ctypes.data_as(fp))
    d1 = np.asarray([1 if (d == 0 and order >= 1) else 0 for d in range(dim)], dtype=np.int32)
    L.intp_ref32_deriv(h, pts.ctypes.data_as(fp), C.c_uint64(len(pts)), d1.ctypes.data_as(C.POINTER(C.c_int)), dv.ctypes.data_as(fp))
    L.intp_ref32_spans(h, pts.ctypes.data_as(fp), C.c_uint64(len(pts)), spans.ctypes.data_as(C.POINTER(C.c_int64)))
    ctrl = np.empty(int(np.prod(shape)), dtype=np.float32)
    L.intp_ref32_ctrl(h, ctrl.ctypes.data_as(fp))
    L.intp_ref32_destroy(h)
    out.update({"c%d_order" % c: order, "c%d_periodic" % c: np.array(per), "c%d_f" % c: f, "c%d_lo" % c: lo, "c%d_hi" % c: hi,
                "c%d_pts" % c: pts, "c%d_vals" % c: vals, "c%d_d1" % c: d1, "c%d_dvals" % c: dv, "c%d_spans" % c: spans,
                "c%d_ctrl" % c: ctrl.reshape(shape)})
out["n_cases"] = len(cases)
np.savez_compressed(os.path.join(HERE, "ref_outputs_f32.npz"), **out)
print("wrote ref_outputs_f32.npz,", len(cases), "cases")
