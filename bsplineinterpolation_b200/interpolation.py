"""Python mirror of the reference's user API for the hot path, over the C ABI.

Names and argument meaning follow the reference headers:
  InterpolationFunctionTemplate  <- InterpolationTemplate.hpp:32-581
  InterpolationFunction          <- Interpolation.hpp:17-507
  BSpline.from_knots             <- BSpline.hpp:188-210
`<T, D, Order, U>` become run-time (dtype, len(shape), order).  numpy arrays play
the role of intp::Mesh (row-major).  Host arrays (numpy) and device arrays (torch
CUDA tensors) are both accepted for meshes, points and outputs; device arrays
are used in place on the given stream.  Everything numeric happens in the CUDA
library; this file only marshals pointers.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import check, lib

_NP = {_capi.F64: np.float64, _capi.F32: np.float32}


def _dtype_code(dtype):
    dt = np.dtype(dtype) if not isinstance(dtype, int) else None
    if dt is None:
        return dtype
    if dt == np.float64:
        return _capi.F64
    if dt == np.float32:
        return _capi.F32
    raise TypeError("dtype must be float64 or float32")


def _is_device(x):
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _i64(v):
    a = np.ascontiguousarray(v, dtype=np.int64)
    return a, a.ctypes.data_as(C.POINTER(C.c_int64))


def _i32(v):
    a = np.ascontiguousarray(v, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(v):
    a = np.ascontiguousarray(v, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def _split_ranges(dim, ranges, shape=None, periodicity=None):
    """ranges[d] is (x_min, x_max) -> uniform axis, or a 1-D coordinate array ->
    non-uniform axis (the value-pair / iterator-pair overloads of the reference).
    A coordinate array must hold shape[d] (+1 on a periodic axis: the closing abscissa)
    values, as the reference asserts (Interpolation.hpp:541); a 2-element list on an axis
    that is not 2 points long is read as (min, max)."""
    if len(ranges) != dim:
        raise ValueError("one range per dimension is required")
    lo, hi, coords = [], [], []
    for d, r in enumerate(ranges):
        arr = np.asarray(r, dtype=np.float64)
        want = None if shape is None else int(shape[d]) + int(bool(periodicity[d]) if periodicity is not None else 0)
        pair = arr.ndim == 1 and arr.size == 2 and (isinstance(r, tuple) or (want is not None and want != 2))
        if pair:
            lo.append(arr[0]); hi.append(arr[1]); coords.append(None)
        elif arr.ndim == 1 and arr.size >= 2:
            if want is not None and arr.size != want:
                raise ValueError("axis %d: coordinate array has %d values, the mesh needs %d%s"
                                 % (d, arr.size, want, " (periodic: closing abscissa included)"
                                    if periodicity is not None and periodicity[d] else ""))
            lo.append(arr[0]); hi.append(arr[-1]); coords.append(np.ascontiguousarray(arr))
        else:
            raise ValueError("range must be a (min, max) tuple or a coordinate array")
    return lo, hi, coords


def _check_device_out(out, dtype, shape, device=None):
    """A caller-supplied CUDA `out` is written through its raw pointer: it must be a contiguous
    tensor of the spline's dtype with exactly the result's element count, on the right device."""
    if not _is_device(out):
        raise TypeError("out must be a CUDA tensor when the points are on the device")
    if out.dtype != dtype or not out.is_contiguous() or out.numel() != int(np.prod(shape)):
        raise ValueError("out must be a contiguous %s CUDA tensor of %s elements" % (dtype, tuple(shape)))
    if device is not None and out.device != device:
        raise ValueError("out lives on %s, the points on %s" % (out.device, device))


class InterpolationFunction:
    """Device-resident spline(s); one per field of the handle."""

    def __init__(self, order=None, f=None, ranges=None, periodicity=None, dtype=np.float64, device=0,
                 _handle=None):
        self._h = None
        if _handle is not None:
            self._h = _handle
        else:
            f_arr = f if _is_device(f) else np.asarray(f)
            shape = tuple(f_arr.shape)
            tmpl = InterpolationFunctionTemplate(order, shape, ranges, periodicity, dtype=dtype, device=device)
            self._h = tmpl.interpolate(f)._steal()
        self._refresh()

    # -- plumbing
    def _steal(self):
        h, self._h = self._h, None
        return h

    def _refresh(self):
        L = lib()
        dt, dim, order = C.c_int(), C.c_int(), C.c_int()
        nf = C.c_int64()
        n = (C.c_int64 * 8)()
        per = (C.c_int * 8)()
        uni = (C.c_int * 8)()
        nk = (C.c_int64 * 8)()
        lo = (C.c_double * 8)()
        hi = (C.c_double * 8)()
        check(L.bspl_function_info(self._h, dt, dim, order, nf, n, per, uni, nk, lo, hi))
        self.dtype = _NP[dt.value]
        self.dim, self.order, self.n_fields = dim.value, order.value, nf.value
        self.shape = tuple(n[d] for d in range(self.dim))
        self._periodic = [bool(per[d]) for d in range(self.dim)]
        self._uniform = [bool(uni[d]) for d in range(self.dim)]
        self._n_knots = [nk[d] for d in range(self.dim)]
        self._range = [(lo[d], hi[d]) for d in range(self.dim)]
        dev = C.c_int(-1)
        check(L.bspl_function_device(self._h, C.byref(dev)))
        self.device = dev.value

    def _check_device(self, t):
        if t.device.index != self.device:
            raise ValueError("tensor on %s, the spline lives on cuda:%d" % (t.device, self.device))

    def __del__(self):
        try:
            if self._h:
                lib().bspl_function_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def copy(self):
        out = C.c_void_p()
        check(lib().bspl_function_clone(self._h, C.byref(out)))
        return InterpolationFunction(_handle=out)

    # -- properties (Interpolation.hpp:248-267)
    def periodicity(self, d):
        return self._periodic[d]

    def uniform(self, d):
        return self._uniform[d]

    def range(self, d):
        return self._range[d]

    def get_order(self):
        return self.order

    def knots(self, d):
        out = np.empty(self._n_knots[d])
        check(lib().bspl_function_knots(self._h, d, out.ctypes.data_as(C.POINTER(C.c_double)), out.size))
        return out

    def control_points(self, field=0):
        out = np.empty(self.shape, dtype=self.dtype)
        check(lib().bspl_function_control_points(self._h, field, out.ctypes.data_as(C.c_void_p)))
        return out

    # -- evaluation
    def _marshal(self, points, out, n_out, fields=1, stream=None):
        if _is_device(points):
            import torch
            pts = points.contiguous()
            if pts.dtype != (torch.float64 if self.dtype == np.float64 else torch.float32):
                raise TypeError("device points must have the spline's dtype")
            self._check_device(pts)
            q = pts.numel() // self.dim
            shape = ((fields,) if fields > 1 else ()) + ((q, n_out) if n_out > 1 else (q,))
            if out is None:
                out = torch.empty(shape, dtype=pts.dtype, device=pts.device)
            else:
                _check_device_out(out, pts.dtype, shape, pts.device)
            sp = stream if stream is not None else torch.cuda.current_stream(pts.device).cuda_stream
            return pts, out, q, C.c_void_p(pts.data_ptr()), C.c_void_p(out.data_ptr()), 1, C.c_void_p(sp)
        pts = np.ascontiguousarray(points, dtype=self.dtype).reshape(-1, self.dim)
        q = pts.shape[0]
        shape = ((fields,) if fields > 1 else ()) + ((q, n_out) if n_out > 1 else (q,))
        if out is None:
            out = np.empty(shape, dtype=self.dtype)
        elif out.dtype != self.dtype or not out.flags.c_contiguous or out.size != int(np.prod(shape)):
            raise ValueError("out must be a C-contiguous %s array of %s elements" % (self.dtype, shape))
        return pts, out, q, pts.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), 0, None

    def evaluate(self, points, out=None, derivatives=None, field=0, stream=None):
        """Batched operator() (derivatives=None) or derivative(coord, derivatives)."""
        pts, out, q, pp, op, dev, sp = self._marshal(points, out, 1, stream=stream)
        dv = None
        if derivatives is not None:
            if len(derivatives) != self.dim:
                raise ValueError("one derivative order per dimension")
            _keep, dv = _i32(derivatives)
        check(lib().bspl_evaluate(self._h, field, pp, q, dv, op, dev, sp))
        return out

    def __call__(self, *coords):
        """f(x, y, ...) with scalars, or f(points[q][dim])."""
        if len(coords) == self.dim and all(np.isscalar(c) for c in coords):
            return float(self.evaluate(np.asarray(coords, dtype=self.dtype)[None, :])[0])
        (points,) = coords
        return self.evaluate(points)

    def at(self, points, field=0):
        """operator() with the bounds check of Interpolation.hpp:153-169 (raises ValueError)."""
        return self.derivative_at(points, None, field)

    def derivative(self, points, derivatives, field=0):
        return self.evaluate(points, derivatives=derivatives, field=field)

    def derivative_at(self, points, derivatives, field=0):
        pts, out, q, pp, op, dev, sp = self._marshal(np.asarray(points), None, 1)
        dv = None
        if derivatives is not None:
            _keep, dv = _i32(derivatives)
        bad = C.c_int64(-1)
        check(lib().bspl_evaluate_at(self._h, field, pp, q, dv, op, C.byref(bad)))
        return out

    def value_grad(self, points, out=None, field=0, stream=None):
        """Fused value + gradient: [q][1+dim]."""
        pts, out, q, pp, op, dev, sp = self._marshal(points, out, self.dim + 1, stream=stream)
        check(lib().bspl_evaluate_value_grad(self._h, field, pp, q, op, dev, sp))
        return out

    def evaluate_fields(self, points, out=None, stream=None, layout="field_major", derivatives=None):
        """One query set on every field.  layout "field_major": [n_fields][q], as if every field had been
        evaluated on its own; "query_major": [q][n_fields], every query's values side by side (what the
        reference returns for a vector-valued T) -- the layout the many-field kernel produces directly."""
        if layout not in ("field_major", "query_major"):
            raise ValueError("layout must be 'field_major' or 'query_major'")
        pts, out, q, pp, op, dev, sp = self._marshal(points, out, 1, fields=self.n_fields, stream=stream)
        if layout == "query_major":
            dv = None
            if derivatives is not None:
                if len(derivatives) != self.dim:
                    raise ValueError("one derivative order per dimension")
                _keep, dv = _i32(derivatives)
            check(lib().bspl_evaluate_fields_query_major(self._h, pp, q, dv, op, dev, sp))
            return out.reshape(q, self.n_fields) if not _is_device(out) else out.view(q, self.n_fields)
        if derivatives is not None:
            raise ValueError("derivatives of all fields at once: use layout='query_major'")
        check(lib().bspl_evaluate_fields(self._h, pp, q, op, dev, sp))
        return out.reshape(self.n_fields, q) if not _is_device(out) else out.view(self.n_fields, q)

    def eval_proxy(self, points, stream=None):
        """Locate (and, for large 3-D batches, tile-sort) the query points once; the returned
        plan evaluates any function of the same template at them (InterpolationTemplate.hpp:145-176)."""
        return QueryPlan(self, points, stream)

    def locate(self, points):
        """span - order per axis (first control point index), int32 [q][dim]."""
        pts = np.ascontiguousarray(points, dtype=self.dtype).reshape(-1, self.dim)
        cell = np.empty(pts.shape, dtype=np.int32)
        check(lib().bspl_locate(self._h, pts.ctypes.data_as(C.c_void_p), pts.shape[0],
                                cell.ctypes.data_as(C.POINTER(C.c_int32)), 0, None))
        return cell


class QueryPlan:
    """Device analogue of the reference's eval_proxy closure: query-dependent work done once."""

    def __init__(self, fn, points, stream=None):
        self._h = None
        self.dim, self.dtype, self.device = fn.dim, fn.dtype, fn.device
        if _is_device(points):
            import torch
            pts = points.contiguous()
            if pts.device.index != fn.device:
                raise ValueError("points on %s, the spline lives on cuda:%d" % (pts.device, fn.device))
            self.q = pts.numel() // fn.dim
            sp = stream if stream is not None else torch.cuda.current_stream(pts.device).cuda_stream
            ptr, dev, sp = C.c_void_p(pts.data_ptr()), 1, C.c_void_p(sp)
        else:
            pts = np.ascontiguousarray(points, dtype=fn.dtype).reshape(-1, fn.dim)
            self.q = pts.shape[0]
            ptr, dev, sp = pts.ctypes.data_as(C.c_void_p), 0, None
        out = C.c_void_p()
        # a plan depends on the knots only: a template makes one before any field exists
        create = (lib().bspl_template_query_plan_create if isinstance(fn, InterpolationFunctionTemplate)
                  else lib().bspl_query_plan_create)
        check(create(fn._h, ptr, self.q, dev, sp, C.byref(out)))
        self._h = out

    def __call__(self, fn, field=0, derivatives=None, value_grad=False, out=None, device_out=False, stream=None):
        """Evaluate `fn` (same template) at the planned points; results in the original order."""
        n_out = fn.dim + 1 if value_grad else 1
        shape = (self.q, n_out) if value_grad else (self.q,)
        dv = None
        if derivatives is not None:
            _keep, dv = _i32(derivatives)
        if device_out or _is_device(out):
            import torch
            tdt = torch.float64 if self.dtype == np.float64 else torch.float32
            if out is None:
                out = torch.empty(shape, dtype=tdt, device=torch.device("cuda", self.device))
            else:
                _check_device_out(out, tdt, shape, torch.device("cuda", self.device))
            sp = stream if stream is not None else torch.cuda.current_stream(out.device).cuda_stream
            check(lib().bspl_query_plan_evaluate(self._h, fn._h, field, dv, int(value_grad),
                                                 C.c_void_p(out.data_ptr()), 1, C.c_void_p(sp)))
            return out
        if out is None:
            out = np.empty(shape, dtype=self.dtype)
        elif (not isinstance(out, np.ndarray) or out.dtype != self.dtype or not out.flags.c_contiguous
              or out.size != int(np.prod(shape))):
            raise ValueError("out must be a C-contiguous %s array of %s elements" % (self.dtype, shape))
        check(lib().bspl_query_plan_evaluate(self._h, fn._h, field, dv, int(value_grad),
                                             out.ctypes.data_as(C.c_void_p), 0, None))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().bspl_query_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass


class InterpolationFunctionTemplate:
    """Knots + factored collocation matrices of one mesh; interpolate() any number of fields."""

    def __init__(self, order, shape, ranges, periodicity=None, dtype=np.float64, device=0):
        self._h = None
        shape = tuple(int(s) for s in (shape if np.ndim(shape) else (shape,)))
        dim = len(shape)
        if periodicity is None:
            periodicity = [False] * dim
        if np.ndim(periodicity) == 0:
            periodicity = [bool(periodicity)] * dim
        lo, hi, coords = _split_ranges(dim, ranges, shape, periodicity)
        self.dim, self.order, self.shape = dim, int(order), shape
        self.dtype = _NP[_dtype_code(dtype)]
        self.device = device
        _kn, n_p = _i64(shape)
        _kp, per_p = _i32([int(bool(p)) for p in periodicity])
        _kl, lo_p = _f64(lo)
        _kh, hi_p = _f64(hi)
        cp = (C.POINTER(C.c_double) * dim)()
        for d in range(dim):
            cp[d] = coords[d].ctypes.data_as(C.POINTER(C.c_double)) if coords[d] is not None else None
        out = C.c_void_p()
        check(lib().bspl_template_create(_dtype_code(dtype), dim, int(order), n_p, per_p, lo_p, hi_p, cp,
                                         device, C.byref(out)))
        self._h = out

    def __del__(self):
        try:
            if self._h:
                lib().bspl_template_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _mesh_ptr(self, f):
        per = int(np.prod(self.shape))
        if _is_device(f):
            import torch
            want = torch.float64 if self.dtype == np.float64 else torch.float32
            if f.dtype != want or not f.is_contiguous():
                raise TypeError("device mesh must be contiguous and of the template's dtype")
            if f.numel() % per or tuple(f.shape[-self.dim:]) != self.shape:
                raise ValueError("mesh shape %s does not match template %s" % (tuple(f.shape), self.shape))
            sp = torch.cuda.current_stream(f.device).cuda_stream
            return f, C.c_void_p(f.data_ptr()), f.numel() // per, 1, C.c_void_p(sp)
        arr = np.ascontiguousarray(f, dtype=self.dtype)
        if arr.size % per or tuple(arr.shape[-self.dim:]) != self.shape:
            raise ValueError("mesh shape %s does not match template %s" % (arr.shape, self.shape))
        return arr, arr.ctypes.data_as(C.c_void_p), arr.size // per, 0, None

    def interpolate(self, f, into=None):
        """interpolate(mesh) -> new function; interpolate(mesh, into=fn) reuses fn's storage.
        A leading extra axis of `f` is a batch of fields."""
        keep, ptr, n_fields, dev, sp = self._mesh_ptr(f)
        if into is not None:
            check(lib().bspl_template_interpolate_into(self._h, into._h, ptr, n_fields, dev, sp))
            into._refresh()
            return into
        out = C.c_void_p()
        check(lib().bspl_template_interpolate(self._h, ptr, n_fields, dev, sp, C.byref(out)))
        return InterpolationFunction(_handle=out)

    def eval_proxy(self, points, stream=None):
        """InterpolationFunctionTemplate::eval_proxy (InterpolationTemplate.hpp:145-165): the
        query-dependent work done once, before any field is interpolated; the returned plan
        evaluates every function this template produces."""
        return QueryPlan(self, points, stream)

    def axis_info(self, axis):
        """(half bandwidth, cyclic, built on the device) of the solver of one axis; see bspl_template_axis_info."""
        band, cyc, dev = C.c_int(), C.c_int(), C.c_int()
        check(lib().bspl_template_axis_info(self._h, int(axis), C.byref(band), C.byref(cyc), C.byref(dev)))
        return band.value, bool(cyc.value), bool(dev.value)

    def sweep_axis(self, axis, data, outer_sizes, outer_strides, line_stride, stream=None):
        """In-place banded/cyclic solve of template axis `axis` along every line of a device
        tensor (one stage of the separable solve); see bspl_template_sweep_axis."""
        import torch
        assert _is_device(data)
        _a, m = _i64(list(outer_sizes))
        _b, ms = _i64(list(outer_strides))
        sp = stream if stream is not None else torch.cuda.current_stream(data.device).cuda_stream
        check(lib().bspl_template_sweep_axis(self._h, int(axis), C.c_void_p(data.data_ptr()), m, ms,
                                             int(line_stride), C.c_void_p(sp)))
        return data

    def sweep_axis_exchange(self, axis, data, outer_sizes, outer_strides, line_stride, split, peers,
                            peer_devices, peer_outer_strides, peer_line_strides, stream=None):
        """sweep_axis whose backward pass stores each solved row into the (peer-mapped) tensor of
        the rank that owns it; see bspl_template_sweep_axis_exchange.  `peers[r]` is a CUDA
        tensor view whose first element is where this rank's block starts in rank r's buffer."""
        import torch
        n_ranks = len(peers)
        _a, m = _i64(list(outer_sizes))
        _b, ms = _i64(list(outer_strides))
        _c, sp_ = _i64(list(split))
        bases = (C.c_void_p * n_ranks)(*[C.c_void_p(p.data_ptr()) for p in peers])
        _d, devs = _i32(list(peer_devices))
        _e, pms = _i64([v for row in peer_outer_strides for v in row])
        _f, pls = _i64(list(peer_line_strides))
        sp = stream if stream is not None else torch.cuda.current_stream(data.device).cuda_stream
        check(lib().bspl_template_sweep_axis_exchange(self._h, int(axis), C.c_void_p(data.data_ptr()), m, ms,
                                                      int(line_stride), n_ranks, sp_, bases, devs, pms, pls,
                                                      C.c_void_p(sp)))

    def function_from_control_points(self, ctrl):
        """Wrap solved plain control points (numpy or CUDA tensor, leading field axis optional)."""
        keep, ptr, n_fields, dev, sp = self._mesh_ptr(ctrl)
        out = C.c_void_p()
        check(lib().bspl_template_function_from_ctrl(self._h, ptr, n_fields, dev, sp, C.byref(out)))
        return InterpolationFunction(_handle=out)


def InterpolationFunction1D(f, order=3, x_range=None, periodicity=False, dtype=np.float64, device=0):
    """1-D convenience (Interpolation.hpp:509-540): default x range [0, N-1], or [0, N] when
    periodic (INTP_PERIODIC_NO_DUMMY_POINT)."""
    f = np.asarray(f) if not _is_device(f) else f
    n = f.shape[0]
    if x_range is None:
        x_range = (0.0, float(n - (0 if periodicity else 1)))
    return InterpolationFunction(order, f, [x_range], [bool(periodicity)], dtype=dtype, device=device)


def InterpolationFunctionTemplate1D(f_length, order=3, x_range=None, periodicity=False, dtype=np.float64, device=0):
    """1-D template (InterpolationTemplate.hpp:583-604): default x range [0, f_length - 1]."""
    if x_range is None:
        x_range = (0.0, float(f_length - 1))
    return InterpolationFunctionTemplate(order, (int(f_length),), [x_range], [bool(periodicity)], dtype=dtype,
                                         device=device)


class BSpline:
    @staticmethod
    def from_knots(order, periodicity, knots, control_points, dtype=np.float64, device=0):
        """BSpline(periodicity, ctrl_pts, knot ranges...) -> evaluable function handle."""
        ctrl = np.ascontiguousarray(control_points, dtype=_NP[_dtype_code(dtype)])
        dim = ctrl.ndim
        ks = [np.ascontiguousarray(k, dtype=np.float64) for k in knots]
        kp = (C.POINTER(C.c_double) * dim)(*[k.ctypes.data_as(C.POINTER(C.c_double)) for k in ks])
        _a, nk = _i64([len(k) for k in ks])
        _b, nc = _i64(ctrl.shape)
        _c, per = _i32([int(bool(p)) for p in periodicity])
        out = C.c_void_p()
        check(lib().bspl_function_from_control_points(_dtype_code(dtype), dim, int(order), nc, per, kp, nk,
                                                      ctrl.ctypes.data_as(C.c_void_p), 1, device, C.byref(out)))
        return InterpolationFunction(_handle=out)


def bspline(order, is_periodic, ranges, mesh, coords, derivative=None, dtype=np.float64, device=0):
    """The flat one-shot call of the reference's MATLAB wrapper (matlab/bspline.cpp:70-141, Example.m):
    (order, is periodic[D], range[D][2], mesh, coords[Q][D], derivative[D]) -> result[Q].  Builds the
    function and evaluates it (or the requested mixed partial derivative) at every coordinate.  A periodic
    axis takes its samples without the closing one, as the wrapper does (bspline.cpp:83-85)."""
    mesh = np.asarray(mesh)
    dim = mesh.ndim
    fn = InterpolationFunction(int(order), mesh, [tuple(r) for r in np.asarray(ranges, dtype=np.float64).reshape(dim, 2)],
                               [bool(p) for p in np.ravel(is_periodic)], dtype=dtype, device=device)
    dv = None if derivative is None else [int(d) for d in np.ravel(derivative)]
    return fn.evaluate(np.asarray(coords).reshape(-1, dim), derivatives=dv)


def band_solve(a, rhs, p, q, cyclic, device=0):
    """BandLU factor + solve on the device (band-matrix-and-solver-test.cpp shape)."""
    a_arr, a_p = _f64(a)
    x = np.array(rhs, dtype=np.float64, copy=True, order="C")
    n = a_arr.shape[0]
    n_rhs = x.size // n
    check(lib().bspl_band_solve(n, p, q, int(bool(cyclic)), a_p, x.ctypes.data_as(C.POINTER(C.c_double)),
                                n_rhs, device))
    return x


def band_solve_rows(rows, rhs, p, q, cyclic, device=0):
    """The same solver fed with the band itself: rows[n][p+q+1], rows[i][k] = A(i, i + k - p), column
    indices wrapping modulo n on a cyclic matrix (bspl_band_solve_rows)."""
    r_arr, r_p = _f64(rows)
    x = np.array(rhs, dtype=np.float64, copy=True, order="C")
    n = r_arr.shape[0]
    if r_arr.ndim != 2 or r_arr.shape[1] != p + q + 1:
        raise ValueError("rows must be [n][p+q+1]")
    check(lib().bspl_band_solve_rows(n, p, q, int(bool(cyclic)), r_p, x.ctypes.data_as(C.POINTER(C.c_double)),
                                     x.size // n, device))
    return x


def launch_count():
    return lib().bspl_launch_count()


def reset_launch_count():
    lib().bspl_reset_launch_count()


def last_kernel_ms():
    return lib().bspl_last_kernel_ms()


def set_sweep_path(path):
    """Control-point solve: "auto", "lines" (thread-per-line sweeps only) or "tiled" (the L2-resident TMA
    sweep wherever it can address the lines); see bspl_set_sweep_path."""
    check(lib().bspl_set_sweep_path({"auto": 0, "lines": 1, "tiled": 2}.get(path, path)))


def set_fields_path(path):
    """Many-field evaluation: "auto", "gather" (per-query gather out of shared memory) or "contract"
    (cell-sorted contraction)."""
    check(lib().bspl_set_fields_path({"auto": 0, "gather": 1, "contract": 2}.get(path, path)))


def set_eval_path(path):
    check(lib().bspl_set_eval_path({"auto": 0, "direct": 1, "binned": 2}.get(path, path)))
