"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU oracle.

Two checkers live here, neither is ever on the product path:
  * ``OracleSpline``  -- oracle/bspl_oracle.c, the plain-C restatement ("port").
  * ``RefSpline``     -- oracle/_ref/libintp_ref_{cell,plain}.so, the UNMODIFIED
                         reference headers compiled by oracle/Makefile
                         ("reference").  Present whenever `make -f oracle/Makefile`
                         ran in a container that has /root/reference; the built
                         .so files travel to the GPU box.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_u64p = C.POINTER(C.c_uint64)
_ip = C.POINTER(C.c_int)


def build(ref="/root/reference", force=False):
    """Compile the C port and, if the reference checkout exists, the ref shim."""
    need = force or not os.path.exists(os.path.join(OUT, "libbspl_oracle.so"))
    if os.path.isdir(os.path.join(ref, "src/include")):
        for n in ("libintp_ref_cell.so", "libintp_ref_plain.so", "libintp_ref_plain_mt.so"):
            need = need or not os.path.exists(os.path.join(OUT, n))
    if need:
        args = ["make", "-j4", "-f", os.path.join(HERE, "Makefile"), "all", "REF=" + ref]
        if force:
            args.insert(1, "-B")
        subprocess.check_call(args, cwd=HERE)
    # the reference's own test programs against the drop-in headers (needs libbspline_b200.so; make
    # tracks the header dependencies, so this is a no-op when nothing changed)
    if os.path.isdir(os.path.join(ref, "test/src")):
        subprocess.check_call(["make", "-s", "-j4", "-f", os.path.join(HERE, "Makefile"), "reftests", "REF=" + ref],
                              cwd=HERE, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, t=_dp):
    return a.ctypes.data_as(t)


# --------------------------------------------------------------------------- port
_port = None


def port_lib():
    global _port
    if _port is None:
        path = os.path.join(OUT, "libbspl_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.bsplo_create.restype = C.c_void_p
        L.bsplo_create.argtypes = [C.c_int, C.c_int, _i64p, _ip, _dp, _dp, C.POINTER(_dp), C.c_int]
        L.bsplo_from_knots.restype = C.c_void_p
        L.bsplo_from_knots.argtypes = [C.c_int, C.c_int, _i64p, _ip, C.POINTER(_dp), _i64p, _dp]
        L.bsplo_destroy.argtypes = [C.c_void_p]
        L.bsplo_interpolate.argtypes = [C.c_void_p, _dp, C.c_int]
        L.bsplo_spans.argtypes = [C.c_void_p, _dp, C.c_int64, _i64p]
        L.bsplo_eval.argtypes = [C.c_void_p, _dp, C.c_int64, _dp, C.c_int]
        L.bsplo_deriv.argtypes = [C.c_void_p, _dp, C.c_int64, _ip, _dp, C.c_int]
        L.bsplo_band_solve.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_int, _dp, _dp]
        _port = L
    return _port


class _Axis(C.Structure):
    _fields_ = [("order", C.c_int), ("periodic", C.c_int), ("uniform", C.c_int),
                ("n", C.c_int64), ("K", C.c_int64), ("t", _dp), ("first", C.c_double),
                ("second", C.c_double), ("dx", C.c_double), ("coords", _dp),
                ("p", C.c_int64), ("q", C.c_int64), ("band", _dp), ("right", _dp),
                ("bottom", _dp)]


class _Spline(C.Structure):
    _fields_ = [("dim", C.c_int), ("order", C.c_int), ("ax", _Axis * 4), ("ctrl", _dp),
                ("size", C.c_int64)]


def _coord_ptrs(dim, coords):
    arr = (_dp * dim)()
    keep = []
    for d in range(dim):
        if coords is not None and coords[d] is not None:
            a = _f64(coords[d])
            keep.append(a)
            arr[d] = _ptr(a)
        else:
            arr[d] = None
    return arr, keep


class OracleSpline:
    """The C restatement. ``f`` given -> interpolate; else knots-only."""

    def __init__(self, order, shape, periodic, lo=None, hi=None, coords=None, f=None,
                 nthreads=1, _handle=None):
        L = port_lib()
        self.L = L
        if _handle is not None:
            self.h = _handle
        else:
            dim = len(shape)
            n = np.asarray(shape, dtype=np.int64)
            per = np.asarray([int(bool(p)) for p in periodic], dtype=np.int32)
            lo_a = _f64(lo if lo is not None else [0.0] * dim)
            hi_a = _f64(hi if hi is not None else [1.0] * dim)
            cp, keep = _coord_ptrs(dim, coords)
            self.h = L.bsplo_create(dim, order, _ptr(n, _i64p), _ptr(per, _ip), _ptr(lo_a),
                                    _ptr(hi_a), cp, 1)
            if not self.h:
                raise ValueError("bsplo_create failed")
        self.s = C.cast(self.h, C.POINTER(_Spline)).contents
        self.dim, self.order = self.s.dim, self.s.order
        self.shape = tuple(int(self.s.ax[d].n) for d in range(self.dim))
        if f is not None:
            self.interpolate(f, nthreads)

    @classmethod
    def from_knots(cls, order, periodic, knots, ctrl):
        L = port_lib()
        ctrl = _f64(ctrl)
        dim = ctrl.ndim
        ks = [_f64(k) for k in knots]
        kp = (_dp * dim)(*[_ptr(k) for k in ks])
        nk = np.asarray([len(k) for k in ks], dtype=np.int64)
        nc = np.asarray(ctrl.shape, dtype=np.int64)
        per = np.asarray([int(bool(p)) for p in periodic], dtype=np.int32)
        h = L.bsplo_from_knots(dim, order, _ptr(nc, _i64p), _ptr(per, _ip), kp, _ptr(nk, _i64p),
                               _ptr(ctrl))
        return cls(order, ctrl.shape, periodic, _handle=h)

    def interpolate(self, f, nthreads=1):
        f = _f64(f)
        assert f.shape == self.shape, (f.shape, self.shape)
        rc = self.L.bsplo_interpolate(self.h, _ptr(f), nthreads)
        assert rc == 0
        return self

    def knots(self, d):
        a = self.s.ax[d]
        return np.ctypeslib.as_array(a.t, shape=(a.K,)).copy()

    def range(self, d):
        return (self.s.ax[d].first, self.s.ax[d].second)

    def control_points(self):
        return np.ctypeslib.as_array(self.s.ctrl, shape=(self.s.size,)).reshape(self.shape).copy()

    def lu(self, d):
        a = self.s.ax[d]
        w = 1 + a.p + a.q
        band = np.ctypeslib.as_array(a.band, shape=(a.n, w)).copy()
        right = bottom = None
        if a.periodic and a.p > 0:
            right = np.ctypeslib.as_array(a.right, shape=(a.n - a.q - 1, a.p)).copy()
            bottom = np.ctypeslib.as_array(a.bottom, shape=(a.n - a.p - 1, a.q)).copy()
        return band, right, bottom

    def spans(self, pts):
        pts = _f64(pts).reshape(-1, self.dim)
        out = np.empty(pts.shape, dtype=np.int64)
        self.L.bsplo_spans(self.h, _ptr(pts), len(pts), _ptr(out, _i64p))
        return out

    def eval(self, pts, nthreads=1):
        pts = _f64(pts).reshape(-1, self.dim)
        out = np.empty(len(pts))
        self.L.bsplo_eval(self.h, _ptr(pts), len(pts), _ptr(out), nthreads)
        return out

    def deriv(self, pts, d, nthreads=1):
        pts = _f64(pts).reshape(-1, self.dim)
        dv = np.asarray(d, dtype=np.int32)
        out = np.empty(len(pts))
        self.L.bsplo_deriv(self.h, _ptr(pts), len(pts), _ptr(dv, _ip), _ptr(out), nthreads)
        return out

    def __del__(self):
        try:
            if self.h:
                self.L.bsplo_destroy(self.h)
                self.h = None
        except Exception:
            pass


def port_band_solve(a, rhs, p, q, cyclic):
    a = _f64(a)
    x = _f64(rhs).copy()
    port_lib().bsplo_band_solve(a.shape[0], p, q, int(cyclic), _ptr(a), _ptr(x))
    return x


# ---------------------------------------------------------------------- reference
_ref = {}


def ref_available():
    return all(os.path.exists(os.path.join(OUT, n))
               for n in ("libintp_ref_cell.so", "libintp_ref_plain.so"))


def ref_lib(kind):
    """kind: 'cell' (INTP_CELL_LAYOUT + INTP_MULTITHREAD), 'plain', or 'plain_mt' (plain layout +
    INTP_MULTITHREAD: the reference's threaded solve without the cell-layout fill)."""
    if kind not in _ref:
        L = C.CDLL(os.path.join(OUT, "libintp_ref_%s.so" % kind))
        L.intp_ref_create.restype = C.c_void_p
        L.intp_ref_create.argtypes = [C.c_int, C.c_int, _u64p, _ip, _dp, _dp, C.POINTER(_dp),
                                      _u64p, _dp, C.c_int]
        L.intp_ref_destroy.argtypes = [C.c_void_p]
        L.intp_ref_eval.argtypes = [C.c_void_p, _dp, C.c_uint64, _dp, C.c_int]
        L.intp_ref_deriv.argtypes = [C.c_void_p, _dp, C.c_uint64, _ip, _dp, C.c_int]
        L.intp_ref_spans.argtypes = [C.c_void_p, _dp, C.c_uint64, _i64p]
        L.intp_ref_knots.restype = C.c_uint64
        L.intp_ref_knots.argtypes = [C.c_void_p, C.c_int, _dp]
        L.intp_ref_range.argtypes = [C.c_void_p, C.c_int, _dp]
        L.intp_ref_ctrl_size.restype = C.c_uint64
        L.intp_ref_ctrl_size.argtypes = [C.c_void_p]
        L.intp_ref_ctrl.argtypes = [C.c_void_p, _dp]
        L.intp_ref_time_template_ms.restype = C.c_double
        L.intp_ref_time_template_ms.argtypes = [C.c_void_p]
        L.intp_ref_time_interpolate_ms.restype = C.c_double
        L.intp_ref_time_interpolate_ms.argtypes = [C.c_void_p]
        L.intp_ref_band_solve.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, _dp, _dp]
        _ref[kind] = L
    return _ref[kind]


class RefSpline:
    """intp::InterpolationFunction<double, D, O> built by the reference itself."""

    def __init__(self, order, f, periodic, lo=None, hi=None, coords=None, kind="cell", repeat=1):
        self.L = ref_lib(kind)
        self.kind = kind
        f = _f64(f)
        self.dim, self.order, self.shape = f.ndim, order, f.shape
        dim = f.ndim
        n = np.asarray(f.shape, dtype=np.uint64)
        per = np.asarray([int(bool(p)) for p in periodic], dtype=np.int32)
        lo_a = _f64(lo if lo is not None else [0.0] * dim)
        hi_a = _f64(hi if hi is not None else [1.0] * dim)
        cp, keep = _coord_ptrs(dim, coords)
        nco = np.asarray([0 if (coords is None or coords[d] is None) else len(coords[d])
                          for d in range(dim)], dtype=np.uint64)
        self.h = self.L.intp_ref_create(dim, order, _ptr(n, _u64p), _ptr(per, _ip), _ptr(lo_a),
                                        _ptr(hi_a), cp, _ptr(nco, _u64p), _ptr(f), repeat)
        if not self.h:
            raise ValueError("intp_ref_create failed (unsupported combination?)")

    def knots(self, d):
        k = self.L.intp_ref_knots(self.h, d, None)
        out = np.empty(k)
        self.L.intp_ref_knots(self.h, d, _ptr(out))
        return out

    def range(self, d):
        out = np.empty(2)
        self.L.intp_ref_range(self.h, d, _ptr(out))
        return (out[0], out[1])

    def control_points(self):
        assert self.kind in ("plain", "plain_mt"), "plain control points need a non-cell-layout build"
        out = np.empty(self.L.intp_ref_ctrl_size(self.h))
        self.L.intp_ref_ctrl(self.h, _ptr(out))
        return out.reshape(self.shape)

    def spans(self, pts):
        pts = _f64(pts).reshape(-1, self.dim)
        out = np.empty(pts.shape, dtype=np.int64)
        self.L.intp_ref_spans(self.h, _ptr(pts), len(pts), _ptr(out, _i64p))
        return out

    def eval(self, pts, nthreads=1):
        pts = _f64(pts).reshape(-1, self.dim)
        out = np.empty(len(pts))
        self.L.intp_ref_eval(self.h, _ptr(pts), len(pts), _ptr(out), nthreads)
        return out

    def deriv(self, pts, d, nthreads=1):
        pts = _f64(pts).reshape(-1, self.dim)
        dv = np.asarray(d, dtype=np.int32)
        out = np.empty(len(pts))
        self.L.intp_ref_deriv(self.h, _ptr(pts), len(pts), _ptr(dv, _ip), _ptr(out), nthreads)
        return out

    @property
    def template_ms(self):
        return self.L.intp_ref_time_template_ms(self.h)

    @property
    def interpolate_ms(self):
        return self.L.intp_ref_time_interpolate_ms(self.h)

    def __del__(self):
        try:
            if self.h:
                self.L.intp_ref_destroy(self.h)
                self.h = None
        except Exception:
            pass


def ref_band_solve(a, rhs, p, q, cyclic, kind="plain"):
    a = _f64(a)
    x = _f64(rhs).copy()
    rc = ref_lib(kind).intp_ref_band_solve(a.shape[0], p, q, int(cyclic), _ptr(a), _ptr(x))
    assert rc == 0
    return x
