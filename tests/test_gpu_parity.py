"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle
(and the reference's golden vectors) on identical inputs.

Tolerances (BASELINE.json north_star): fp64 values / derivatives / control points
<= 1e-12 relative (reference rel_err metric AND max-abs vs field scale); span and
index selection bit-exact.  Control points are in fact expected to be bit-identical
(same elimination order, unfused arithmetic)."""
import numpy as np
import pytest

from conftest import rel_err
from cases import adversarial_points, all_combos, axis_ranges, extended_combos, queries, small_shapes, smooth_field
from oracle.pyoracle import OracleSpline

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _close(a, b, tol=TOL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    assert np.abs(a - b).max() <= tol * scale, (np.abs(a - b).max(), scale)
    if np.abs(b).sum() > 0:
        assert rel_err(a, b) <= tol


@pytest.fixture(scope="module")
def pkg(lib_built):
    return lib_built


def _same_control_points(c, ref):
    """Long axes may take the windowed one-pass sweep (warm-up error 5e-19 relative): equal to
    1e-13 of the scale and bit-identical but for isolated last-bit ties."""
    assert np.abs(c - ref).max() <= 1e-13 * np.abs(ref).max()
    assert (c != ref).mean() < 1e-3, (c != ref).mean()


def _ranges(lo, hi):
    return [(float(a), float(b)) for a, b in zip(lo, hi)]


@pytest.mark.parametrize("dim,order,periodic", list(all_combos()))
def test_solve_and_eval_match_oracle(pkg, dim, order, periodic):
    rng = np.random.default_rng(1000 * dim + 10 * order + sum(periodic))
    shape = small_shapes(dim, order, periodic)
    lo, hi = axis_ranges(dim, rng)
    f = smooth_field(shape, rng)
    o = OracleSpline(order, shape, periodic, lo=lo, hi=hi, f=f)
    fn = pkg.InterpolationFunction(order, f, _ranges(lo, hi), periodic)
    for d in range(dim):
        assert np.array_equal(fn.knots(d), o.knots(d))
        assert fn.range(d) == o.range(d)
        assert fn.periodicity(d) == bool(periodic[d]) and fn.uniform(d)
    ctrl = fn.control_points()
    assert np.array_equal(ctrl, o.control_points()), np.abs(ctrl - o.control_points()).max()

    rlo = np.array([o.range(d)[0] for d in range(dim)]); rhi = np.array([o.range(d)[1] for d in range(dim)])
    adv = adversarial_points([o.knots(d) for d in range(dim)], rlo, rhi, periodic, rng)
    assert np.array_equal(fn.locate(adv), o.spans(adv))
    wild = queries(rlo, rhi, periodic, 4000, rng, mode="wild")
    assert np.array_equal(fn.locate(wild), o.spans(wild))

    pts = queries(rlo, rhi, periodic, 3000, rng)
    _close(fn(pts), o.eval(pts))
    _close(fn(wild), o.eval(wild), 1e-10)  # extrapolation amplifies rounding
    for dv in ([1] + [0] * (dim - 1), [min(order, 2)] * dim, [0] * (dim - 1) + [order]):
        _close(fn.derivative(pts, dv), o.deriv(pts, dv))
    assert not fn.derivative(pts, [order + 1] + [0] * (dim - 1)).any()  # BSpline.hpp:404-407
    vg = fn.value_grad(pts)
    _close(vg[:, 0], o.eval(pts))
    for d in range(dim):
        dv = [0] * dim; dv[d] = 1
        _close(vg[:, 1 + d], o.deriv(pts, dv))


@pytest.mark.parametrize("dim,order,periodic", sorted(set(extended_combos())))
def test_extended_dims_and_orders_match_oracle(pkg, dim, order, periodic):
    """What the reference's templates accept beyond the everyday set (Interpolation.hpp:17 takes any D and Order):
    orders 6 and 7 -- half bandwidth 5 and 6 on non-periodic axes -- and 4-D splines, served by the run-time-loop
    kernels (eval_generic_kernel, the thread-per-line sweep for wide bands, one sweep launch per field in 4-D).
    Control points and spans bit-identical to the oracle (pinned against the reference on these very
    combinations, tests/test_oracle.py); values / derivatives / gradient to 1e-12 of the magnitude of the terms
    summed: on these tiny high-order meshes the control points reach 10^3 - 10^4 times the field (the collocation
    problem is that ill-conditioned), so the bar is 1e-12 * max|control point| / max|field| relative."""
    rng = np.random.default_rng(7000 + 100 * dim + 10 * order + sum(periodic))
    shape = small_shapes(dim, order, periodic)
    lo, hi = axis_ranges(dim, rng)
    f = smooth_field(shape, rng)
    o = OracleSpline(order, shape, periodic, lo=lo, hi=hi, f=f)
    fn = pkg.InterpolationFunction(order, f, _ranges(lo, hi), periodic)
    for d in range(dim):
        assert np.array_equal(fn.knots(d), o.knots(d))
        assert fn.range(d) == o.range(d)
    assert np.array_equal(fn.control_points(), o.control_points())
    tol = 1e-12 * max(1.0, np.abs(o.control_points()).max() / np.abs(f).max())
    rlo = np.array([o.range(d)[0] for d in range(dim)]); rhi = np.array([o.range(d)[1] for d in range(dim)])
    adv = adversarial_points([o.knots(d) for d in range(dim)], rlo, rhi, periodic, rng)
    assert np.array_equal(fn.locate(adv), o.spans(adv))
    pts = queries(rlo, rhi, periodic, 2000, rng)
    assert np.array_equal(fn.locate(pts), o.spans(pts))
    _close(fn(pts), o.eval(pts), tol)
    for dv in ([1] + [0] * (dim - 1), [min(order, 2)] * dim, [0] * (dim - 1) + [order]):
        _close(fn.derivative(pts, dv), o.deriv(pts, dv), tol)
    assert not fn.derivative(pts, [order + 1] + [0] * (dim - 1)).any()
    vg = fn.value_grad(pts)
    _close(vg[:, 0], o.eval(pts), tol)
    for d in range(dim):
        dv = [0] * dim; dv[d] = 1
        _close(vg[:, 1 + d], o.deriv(pts, dv), tol)
    # two fields on one template, device pointers, and the many-field entry points (fallback routes here)
    import torch
    f2 = np.stack([f, smooth_field(shape, rng)])
    t = pkg.InterpolationFunctionTemplate(order, shape, _ranges(lo, hi), periodic)
    fn2 = t.interpolate(torch.from_numpy(f2).cuda())
    o2 = OracleSpline(order, shape, periodic, lo=lo, hi=hi, f=f2[1])
    assert np.array_equal(fn2.control_points(field=1), o2.control_points())
    dp = torch.from_numpy(pts).cuda()
    both = fn2.evaluate_fields(dp).cpu().numpy()
    _close(both[0], o.eval(pts), tol); _close(both[1], o2.eval(pts), 10 * tol)
    qm = fn2.evaluate_fields(dp, layout="query_major").cpu().numpy()
    _close(qm[:, 1], o2.eval(pts), 10 * tol)


def test_reference_golden_vectors(pkg, golden):
    """Every known-answer vector of interpolation-test.cpp that is on the path (tol 1e-14, :16)."""
    g = golden["interpolation"]
    tol = 1e-14
    IF = pkg.InterpolationFunction
    f = np.array(g["f"])
    xs_h = np.array(g["coords_1d_half"]); xs = np.array(g["coords_1d"])
    cubic = IF(3, f, [(0.0, 6.0)])
    assert rel_err(cubic(xs_h), g["vals_1d"]) < tol
    assert rel_err(cubic.derivative_at(xs_h, [1]), g["vals_1d_derivative_1"]) < tol
    assert abs(cubic(-0.5) - g["extrapolate_left"][1]) < tol
    assert abs(cubic(6.5) - g["extrapolate_right"][1]) < tol
    idx = np.floor(xs).astype(int)
    assert rel_err(IF(1, f, [(0.0, 12.0)])(xs), f[idx] + (f[idx + 1] - f[idx]) * (xs - idx)) < tol
    assert rel_err(IF(0, f, [(0.0, 12.0)])(xs), f[np.round(xs).astype(int)]) < tol
    fp = f[:-1]
    quart = IF(4, fp, [(0.0, 12.0)], [True])
    assert rel_err(quart(xs), g["vals_1d_periodic"]) < tol
    assert rel_err(quart.derivative_at(xs, [1]), g["vals_1d_derivative_periodic"]) < tol
    assert abs(quart(xs[0] - 12) - g["vals_1d_periodic"][0]) < tol
    assert abs(quart(xs[0] + 12) - g["vals_1d_periodic"][0]) < tol
    assert rel_err(IF(1, fp, [(0.0, 12.0)], [True])(xs), fp[idx] + (fp[(idx + 1) % 12] - fp[idx]) * (xs - idx)) < tol
    f2 = np.array(g["f2"]).reshape(5, 5); c2 = np.array(g["coords_2d"]).reshape(-1, 2)
    s2 = IF(3, f2, [(0.0, 4.0), (0.0, 4.0)])
    assert s2.uniform(0) and s2.uniform(1)
    assert rel_err(s2(c2), g["vals_2d"]) < tol
    assert rel_err(s2.derivative_at(c2, [2, 1]), g["vals_2d_derivative_x2_y1"]) < tol
    with pytest.raises(ValueError):  # std::domain_error, :159-164
        s2.at(np.array([[-1.0, 1.0]]))
    assert rel_err(IF(3, f2[:, :4], [(0.0, 4.0), (0.0, 4.0)], [False, True])(c2), g["vals_2d_periodic"]) < tol
    f3 = np.array(g["f3"]).reshape(5, 6, 7); c3 = np.array(g["coords_3d"]).reshape(-1, 3)
    s3 = IF(3, f3, [(0.0, 4.0), (0.0, 5.0), (0.0, 6.0)])
    assert rel_err(s3(c3), g["vals_3d"]) < tol
    assert rel_err(s3.derivative_at(c3, [1, 0, 3]), g["vals_3d_derivative_x1_y0_z3"]) < tol
    assert not any(s3.periodicity(d) for d in range(3))
    # non-uniform axes (:517-670)
    xc = np.array(g["input_coords_1d"])
    assert rel_err(IF(3, f, [xc])(xs), g["vals_1d_nonuniform"]) < tol
    nup = IF(4, fp, [xc], [True])
    assert not nup.uniform(0)
    assert rel_err(nup(xs), g["vals_1d_nonuniform_periodic"]) < tol
    mixed = IF(3, f2[:4], [(0.0, 4.0), np.array(g["nonuniform_coord_for_2d"])], [True, False])
    assert rel_err(mixed(c2), g["vals_2d_X_periodic_Y_nonuniform"]) < tol


def test_bspline_golden_vectors(pkg, golden):
    """bspline-test.cpp: spline from knots + control points (tol 1e-15, :45)."""
    b = golden["bspline"]
    tol = 2e-15
    k = b["knots"]
    s1 = pkg.BSpline.from_knots(3, [0], [k], np.array(b["cp"]))
    x1 = np.array(b["coords_1d"])
    assert rel_err(s1(x1), b["vals_1d"]) < tol
    assert rel_err(s1.derivative(x1, [0]), b["vals_1d"]) < tol
    assert rel_err(s1.derivative(x1, [1]), b["vals_1d_derivative_1"]) < tol
    assert rel_err(s1.derivative(x1, [2]), b["vals_1d_derivative_2"]) < tol
    cp2 = np.array(b["cp2"]).reshape(5, 5); x2 = np.array(b["coords_2d"]).reshape(-1, 2)
    s2 = pkg.BSpline.from_knots(3, [0, 0], [k, k], cp2)
    assert rel_err(s2(x2), b["vals_2d"]) < tol
    assert rel_err(s2.derivative(x2, [2, 0]), b["vals_2d_derivative_x2_y0"]) < tol
    assert rel_err(s2.derivative(x2, [1, 1]), b["vals_2d_derivative_x1_y1"]) < tol
    s2p = pkg.BSpline.from_knots(3, [0, 1], [k, b["knots2"]], cp2)
    assert rel_err(s2p(x2), b["vals_2d_periodic"]) < tol
    assert rel_err(s2p.derivative(x2, [1, 1]), b["vals_2d_periodic_derivative_x1_y1"]) < tol
    cp3 = np.array(b["cp3"]).reshape(5, 5, 5); x3 = np.array(b["coords_3d"]).reshape(-1, 3)
    s3 = pkg.BSpline.from_knots(3, [0, 0, 0], [k, k, k], cp3)
    assert rel_err(s3(x3), b["vals_3d"]) < tol


def test_committed_reference_outputs(pkg):
    """Outputs of the unmodified reference (tests/golden/ref_outputs.npz)."""
    import os
    from conftest import ROOT
    data = np.load(os.path.join(ROOT, "tests", "golden", "ref_outputs.npz"))
    for c in range(int(data["n_cases"])):
        order = int(data["c%d_order" % c]); per = [bool(v) for v in data["c%d_periodic" % c]]
        f = data["c%d_f" % c]
        fn = pkg.InterpolationFunction(order, f, _ranges(data["c%d_lo" % c], data["c%d_hi" % c]), per)
        assert np.array_equal(fn.control_points(), data["c%d_ctrl" % c])
        pts = data["c%d_pts" % c]
        assert np.array_equal(fn.locate(pts), data["c%d_spans" % c])
        _close(fn(pts), data["c%d_vals" % c], 1e-10)  # includes extrapolated points
        _close(fn.derivative(pts, [int(v) for v in data["c%d_dv" % c]]), data["c%d_dvals" % c], 1e-9)


def test_committed_reference_outputs_fp32(pkg):
    """The reference instantiated with T = U = float (tests/golden/ref_outputs_f32.npz, made by
    make_ref_outputs_f32.py): spans bit-exact, values / derivatives / control points within the north star's
    fp32 tolerance of 1e-5 (relative to the largest magnitude)."""
    import os
    from conftest import ROOT
    data = np.load(os.path.join(ROOT, "tests", "golden", "ref_outputs_f32.npz"))
    tol = 1e-5
    for c in range(int(data["n_cases"])):
        order = int(data["c%d_order" % c]); per = [bool(v) for v in data["c%d_periodic" % c]]
        f = data["c%d_f" % c]
        fn = pkg.InterpolationFunction(order, f, _ranges(data["c%d_lo" % c], data["c%d_hi" % c]), per, dtype=np.float32)
        assert fn.dtype == np.float32
        ctrl = data["c%d_ctrl" % c]
        assert np.abs(fn.control_points() - ctrl).max() <= tol * np.abs(ctrl).max()
        pts = data["c%d_pts" % c]
        assert pts.dtype == np.float32
        assert np.array_equal(fn.locate(pts), data["c%d_spans" % c])
        vals = data["c%d_vals" % c]
        assert np.abs(fn(pts) - vals).max() <= tol * np.abs(vals).max()
        dvals = data["c%d_dvals" % c]
        got = fn.derivative(pts, [int(v) for v in data["c%d_d1" % c]])
        assert np.abs(got - dvals).max() <= tol * np.abs(dvals).max()


def test_long_axis_compact_factors_match_committed_reference_outputs(pkg):
    """tests/golden/ref_outputs_long.npz (the unmodified reference on a 16 411-point uniform axis): such
    axes are factored in compact form (both ends on a surrogate, one steady row in between,
    bspl_host.h) and solved by the chunk-parallel sweep.  Spans exact, control points to 1e-13
    (bit-identical but for isolated last-bit ties), values to 1e-12."""
    import os
    from cases import long_axis_field
    from conftest import ROOT
    d = np.load(os.path.join(ROOT, "tests", "golden", "ref_outputs_long.npz"))
    n = int(d["n"])
    f = long_axis_field(n)
    for c in range(int(d["n_cases"])):
        order, per = int(d["c%d_order" % c]), bool(d["c%d_periodic" % c])
        fn = pkg.InterpolationFunction(order, f, [(-1.5, 2.25)], [per])
        ctrl, ref = fn.control_points(), d["c%d_ctrl" % c]
        assert np.abs(ctrl - ref).max() <= 1e-13 * np.abs(ref).max(), (order, per)
        assert (ctrl != ref).mean() < 1e-3
        pts = d["c%d_pts" % c]
        assert np.array_equal(fn.locate(pts).reshape(-1), d["c%d_spans" % c].reshape(-1))
        _close(fn(pts), d["c%d_vals" % c], 1e-11 if per else 1e-12)   # wrapped / extrapolated points included
    # the same axis inside a 2-D mesh (48 strided lines of 16 411 points)
    t = pkg.InterpolationFunctionTemplate(3, (n, 48), [(-1.5, 2.25), (0.0, 1.0)], [False, False])
    g = np.outer(f, np.ones(48))   # every column of the mesh is the 1-D data
    fn2 = t.interpolate(g)
    # axis 1 (constant data along a clamped cubic axis) reproduces the constant, so column k equals the 1-D control points
    ref3 = d["c0_ctrl"]
    got = fn2.control_points()
    assert np.abs(got[:, 7] - ref3).max() <= 1e-12 * np.abs(ref3).max()


def test_band_solver(pkg):
    """band-matrix-and-solver-test.cpp: ||b - A x|| / ||b|| < 1e-10, and bit parity with the oracle."""
    from test_oracle import _band_matrices
    from oracle.pyoracle import port_band_solve
    n = 64
    rhs = np.random.default_rng(5).uniform(-1, 1, (3, n))
    for a, p, q, cyc in _band_matrices(n):
        x = pkg.band_solve(a, rhs, p, q, cyc)
        for r in range(3):
            assert np.linalg.norm(a @ x[r] - rhs[r]) / np.linalg.norm(rhs[r]) < 1e-10
            assert np.array_equal(x[r], port_band_solve(a, rhs[r], p, q, cyc))


def test_nonuniform_axes_match_oracle(pkg):
    rng = np.random.default_rng(77)
    for order in range(1, 6):
        for per in (False, True):
            n = 29
            xc = np.sort(rng.uniform(0, 5, n + per)); xc[0] = 0; xc[-1] = 5
            f = rng.standard_normal(n)
            o = OracleSpline(order, (n,), [per], coords=[xc], f=f)
            fn = pkg.InterpolationFunction(order, f, [xc], [per])
            assert np.array_equal(fn.knots(0), o.knots(0))
            assert np.array_equal(fn.control_points(), o.control_points())
            pts = rng.uniform(0, 5, 2000)
            assert np.array_equal(fn.locate(pts)[:, 0], o.spans(pts)[:, 0])
            _close(fn(pts), o.eval(pts))
            _close(fn.derivative(pts, [1]), o.deriv(pts, [1]))


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_long_nonuniform_axis_is_factored_on_the_device(pkg, order, monkeypatch):
    """collocation matrix and band LU of a long non-uniform axis are built on the
    device (bspl_factor.cu: one thread per abscissa for the rows, chunk-parallel elimination with a warm-up
    window, seams checked bit for bit; periodic axes too, with the border of their bordered LU taken from a
    host surrogate of the first pivots).  The solves must equal the oracle's -- whose LU is the reference's
    sequential one -- bit for bit: a 1-D axis of 40 000 points, a 2-D mesh with such an axis, and a short axis
    cut into many small chunks (seams every 64 rows).  Wildly uneven spacing (ratios up to 1e3) is allowed to
    fail the seam check and fall back to the host; the answer is the same either way."""
    rng = np.random.default_rng(8300 + order)
    def coords(n, jitter):
        x = np.arange(n, dtype=np.float64) + rng.uniform(-jitter, jitter, n)
        x[0], x[-1] = 0.0, float(n - 1)
        return np.sort(x) / (n - 1) * 3.0
    n = 40000
    xc = coords(n, 0.35)
    f = rng.standard_normal(n)
    t = pkg.InterpolationFunctionTemplate(order, (n,), [xc], [False])
    assert t.axis_info(0) == (order - 1, False, True)
    o = OracleSpline(order, (n,), [False], coords=[xc], f=f)
    fn = t.interpolate(f)
    assert np.array_equal(fn.knots(0), o.knots(0))
    assert np.array_equal(fn.control_points(), o.control_points())
    pts = rng.uniform(0, 3, 5000)
    _close(fn(pts), o.eval(pts))
    # 2-D, long non-uniform axis first, short uniform periodic axis second
    shape = (20000, 24)
    xc2 = coords(shape[0], 0.3)
    f2 = rng.standard_normal(shape)
    t2 = pkg.InterpolationFunctionTemplate(order, shape, [xc2, (0.0, 1.0)], [False, True])
    assert t2.axis_info(0)[2] and not t2.axis_info(1)[2]
    o2 = OracleSpline(order, shape, [False, True], coords=[xc2, None], lo=[0, 0], hi=[0, 1], f=f2)
    assert np.array_equal(t2.interpolate(f2).control_points(), o2.control_points())
    # many seams on a short axis; then spacing so uneven that the device result may be refused
    monkeypatch.setenv("BSPL_DEVICE_LU_MIN", "200")
    monkeypatch.setenv("BSPL_DEVICE_LU_CHUNK", "64")
    monkeypatch.setenv("BSPL_DEVICE_LU_WINDOW", "96")
    for n3, jitter in ((1500, 0.4), (1500, 0.499)):
        xc3 = coords(n3, jitter)
        f3 = rng.standard_normal(n3)
        t3 = pkg.InterpolationFunctionTemplate(order, (n3,), [xc3], [False])
        if jitter < 0.45:
            assert t3.axis_info(0)[2]
        o3 = OracleSpline(order, (n3,), [False], coords=[xc3], f=f3)
        assert np.array_equal(t3.interpolate(f3).control_points(), o3.control_points())
    monkeypatch.setenv("BSPL_DEVICE_LU_MIN", "0")   # disabled: host path
    t4 = pkg.InterpolationFunctionTemplate(order, (n,), [xc], [False])
    assert not t4.axis_info(0)[2]
    assert np.array_equal(t4.interpolate(f).control_points(), o.control_points())
    monkeypatch.delenv("BSPL_DEVICE_LU_MIN"); monkeypatch.delenv("BSPL_DEVICE_LU_CHUNK"); monkeypatch.delenv("BSPL_DEVICE_LU_WINDOW")
    # periodic: the border of the bordered LU (corner strips, corner block) comes from the host, the band from the device
    if order >= 2:
        xp = coords(n + 1, 0.35)
        tp = pkg.InterpolationFunctionTemplate(order, (n,), [xp], [True])
        assert tp.axis_info(0) == (order // 2, True, True)
        op = OracleSpline(order, (n,), [True], coords=[xp], f=f)
        fp = tp.interpolate(f)
        assert np.array_equal(fp.knots(0), op.knots(0))
        assert np.array_equal(fp.control_points(), op.control_points())
        shape = (24, 20000)
        xp2 = coords(shape[1] + 1, 0.3)
        f2p = rng.standard_normal(shape)
        t2p = pkg.InterpolationFunctionTemplate(order, shape, [(0.0, 1.0), xp2], [False, True])
        assert t2p.axis_info(1)[2]
        o2p = OracleSpline(order, shape, [False, True], coords=[None, xp2], lo=[0, 0], hi=[1, 0], f=f2p)
        assert np.array_equal(t2p.interpolate(f2p).control_points(), o2p.control_points())


def test_template_reuse_many_fields(pkg):
    """InterpolationFunctionTemplate: one mesh, many fields (cfg5 shape, scaled down)."""
    rng = np.random.default_rng(3)
    shape = (32, 24)
    F = 7
    fields = np.stack([smooth_field(shape, rng) for _ in range(F)])
    t = pkg.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0), (-1.0, 1.0)], [True, False])
    fn = t.interpolate(fields)
    assert fn.n_fields == F
    pts = queries(np.array([0.0, -1.0]), np.array([1.0, 1.0]), [True, False], 1500, rng)
    allv = fn.evaluate_fields(pts)
    for k in range(F):
        o = OracleSpline(3, shape, [True, False], lo=[0, -1], hi=[1, 1], f=fields[k])
        assert np.array_equal(fn.control_points(k), o.control_points())
        _close(allv[k], o.eval(pts))
        _close(fn.evaluate(pts, field=k), o.eval(pts))
    # interpolate(fn, mesh): reuse storage
    again = t.interpolate(fields[::-1].copy(), into=fn)
    assert np.array_equal(again.control_points(0), OracleSpline(3, shape, [True, False], lo=[0, -1], hi=[1, 1],
                                                                 f=fields[-1]).control_points())
    cp = fn.copy()
    assert np.array_equal(cp.control_points(2), fn.control_points(2))


def test_device_pointer_entry_points(pkg):
    import torch
    rng = np.random.default_rng(11)
    shape = (40, 36, 28)
    f = smooth_field(shape, rng)
    t = pkg.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 3)
    fn = t.interpolate(torch.from_numpy(f).cuda())
    o = OracleSpline(3, shape, [0, 0, 0], lo=[0, 0, 0], hi=[1, 1, 1], f=f)
    assert np.array_equal(fn.control_points(), o.control_points())
    pts = rng.uniform(0, 1, (5000, 3))
    dv = fn.value_grad(torch.from_numpy(pts).cuda())
    torch.cuda.synchronize()
    vg = dv.cpu().numpy()
    _close(vg[:, 0], o.eval(pts))
    _close(vg[:, 2], o.deriv(pts, [0, 1, 0]))


def test_unaligned_device_output_takes_the_direct_kernel(pkg):
    """The binned kernel stores {value, gradient} as one 32-byte vector; an output pointer that is not
    32-byte aligned must not reach it (launch_eval falls back to the direct kernel)."""
    import torch
    rng = np.random.default_rng(12)
    shape = (32, 32, 32)
    fn = pkg.InterpolationFunction(3, smooth_field(shape, rng), [(0.0, 1.0)] * 3)
    q = 1 << 21
    pts = torch.rand((q, 3), dtype=torch.float64, device="cuda")
    ref = fn.value_grad(pts)                                   # aligned: binned path
    buf = torch.empty(q * 4 + 1, dtype=torch.float64, device="cuda")
    out = buf[1:].view(q, 4)                                   # 8 bytes past a 32-byte boundary
    assert out.data_ptr() % 32 == 8
    fn.value_grad(pts, out=out)
    torch.cuda.synchronize()
    assert (out - ref).abs().max().item() <= 1e-13 * ref.abs().max().item()


def test_empty_and_single_queries(pkg):
    f = np.arange(10.0)
    fn = pkg.InterpolationFunction(3, f, [(0.0, 9.0)])
    assert fn(np.empty((0, 1))).shape == (0,)
    assert abs(fn(4.0) - 4.0) < 1e-12  # cubic spline reproduces a linear function
    assert abs(fn.derivative(np.array([[2.5]]), [1])[0] - 1.0) < 1e-12


def test_medium_3d_properties(pkg):
    """Size-independent checks at a larger size: the interpolant reproduces its data at
    the mesh nodes, and solve is linear."""
    rng = np.random.default_rng(21)
    shape = (96, 80, 72)
    f1 = smooth_field(shape, rng); f2 = smooth_field(shape, rng)
    rngs = [(0.0, 1.0), (0.0, 2.0), (-1.0, 1.0)]
    t = pkg.InterpolationFunctionTemplate(3, shape, rngs, [False, True, False])
    fn = t.interpolate(np.stack([f1, f2, 2.0 * f1 - 3.0 * f2]))
    c = [fn.control_points(k) for k in range(3)]
    assert np.abs(c[2] - (2.0 * c[0] - 3.0 * c[1])).max() <= 1e-12 * np.abs(c[2]).max()
    ax0 = np.linspace(0, 1, shape[0]); ax1 = np.arange(shape[1]) * (2.0 / shape[1]); ax2 = np.linspace(-1, 1, shape[2])
    idx = rng.integers(0, [shape[0], shape[1], shape[2]], size=(4000, 3))
    nodes = np.stack([ax0[idx[:, 0]], ax1[idx[:, 1]], ax2[idx[:, 2]]], axis=1)
    vals = fn.evaluate(nodes, field=0)
    assert np.abs(vals - f1[idx[:, 0], idx[:, 1], idx[:, 2]]).max() <= 1e-12 * np.abs(f1).max()


@pytest.mark.parametrize("order", range(6))
@pytest.mark.parametrize("periodic", [(False, False, False), (True, False, True), (True, True, True)])
def test_binned_tma_path_matches_direct_and_oracle(pkg, order, periodic):
    """The cell-binned / TMA-staged path (forced) must give what the direct gather gives."""
    import torch
    rng = np.random.default_rng(500 + order)
    shape = (45, 38, 52)
    lo, hi = axis_ranges(3, rng)
    f = smooth_field(shape, rng)
    fn = pkg.InterpolationFunction(order, f, _ranges(lo, hi), periodic)
    o = OracleSpline(order, shape, periodic, lo=lo, hi=hi, f=f)
    rlo = np.array([o.range(d)[0] for d in range(3)]); rhi = np.array([o.range(d)[1] for d in range(3)])
    pts = queries(rlo, rhi, periodic, 60000, rng, mode="wild")
    dpts = torch.from_numpy(pts).cuda()
    try:
        pkg.set_eval_path("direct")
        v_d = fn.evaluate(dpts).cpu().numpy(); g_d = fn.value_grad(dpts).cpu().numpy()
        dv = [min(order, 1), 0, min(order, 2)]
        d_d = fn.evaluate(dpts, derivatives=dv).cpu().numpy()
        pkg.set_eval_path("binned")
        v_b = fn.evaluate(dpts).cpu().numpy(); g_b = fn.value_grad(dpts).cpu().numpy()
        d_b = fn.evaluate(dpts, derivatives=dv).cpu().numpy()
        h_b = fn.value_grad(pts)  # host-pointer entry through the same path
    finally:
        pkg.set_eval_path("auto")
    # same weights up to the reciprocal used (interior tiles take a division-free path)
    for b_, d_ in ((v_b, v_d), (g_b, g_d), (d_b, d_d)):
        assert np.abs(b_ - d_).max() <= 1e-13 * max(np.abs(d_).max(), 1e-300)
    assert np.array_equal(h_b, g_b)
    inside = np.all((pts >= rlo) & (pts <= rhi), axis=1) | np.array(periodic).all()
    _close(v_b[inside], o.eval(pts[inside]))
    _close(g_b[inside, 2], o.deriv(pts[inside], [0, 1, 0]))


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
def test_tma_tiled_contiguous_sweep_bit_exact(pkg, order):
    """>= 4096 lines along the contiguous axis take the TMA-tiled sweep (bspl_solve.cu): fused with
    the copy out of the caller's mesh, cyclic shifts of periodic axes applied to the tile
    coordinates (wrapped rows patched in) and along the line (delay of the right-hand side);
    ragged tiles (70 lines per run, 72 = 4.5 tile rows), many fields, host and device meshes, and
    an odd row length that TMA cannot address (falls back).
    Control points bit-identical to the oracle's sequential solve in every case."""
    import torch
    rng = np.random.default_rng(4200 + order)
    cases = [((66, 70, 72), [False, False, False]), ((66, 70, 72), [True, False, False]),
             ((66, 70, 72), [False, True, True]), ((66, 70, 72), [True, True, True]),
             ((66, 70, 72), [False, False, True]), ((66, 70, 72), [False, True, False]),
             ((66, 70, 73), [False, False, False]), ((260, 48), [False, False]), ((260, 48), [True, True]),
             ((130, 40), [False, True])]
    for shape, per in cases:
        dim = len(shape)
        fields = 20 if dim == 2 else 1
        f = rng.standard_normal((fields,) + shape)
        t = pkg.InterpolationFunctionTemplate(order, shape, [(0.0, 1.0 + d) for d in range(dim)], per)
        fn_host = t.interpolate(f if fields > 1 else f[0])
        fn_dev = t.interpolate(torch.from_numpy(f if fields > 1 else f[0]).cuda())
        for k in range(fields):
            o = OracleSpline(order, shape, per, lo=[0.0] * dim, hi=[1.0 + d for d in range(dim)], f=f[k])
            ref = o.control_points()
            assert np.array_equal(fn_host.control_points(field=k), ref), (shape, per, k)
            assert np.array_equal(fn_dev.control_points(field=k), ref), (shape, per, k)
    # float
    f32 = rng.standard_normal((66, 70, 72)).astype(np.float32)
    fn32 = pkg.InterpolationFunction(order, f32, [(0.0, 1.0)] * 3, dtype=np.float32)
    ref = OracleSpline(order, (66, 70, 72), [False] * 3, lo=[0.0] * 3, hi=[1.0] * 3, f=f32.astype(np.float64)).control_points()
    assert np.abs(fn32.control_points() - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("order", [0, 1, 2, 3, 4, 5])
def test_l2_resident_tiled_sweeps_bit_exact(pkg, order):
    """The sweeps that keep their lines in flight inside the L2 (sweep_rows_tma_kernel: TMA row tiles for the
    strided axes, swizzled boxes for the contiguous axis, packed factor rows staged with the tiles, the
    division by the pivot taken off the dependent chain) are the automatic choice only for meshes larger than
    the L2; forced here on small ones: ragged tiles (132 lines = 4.125 warps, 136 = 4.25 tiles of 32 steps),
    periodic axes (corner strips and tail rows), 2-D, many fields, zeros in the data (operands outside the
    fast division's range take the full routine) and float.  Control points bit-identical to the oracle's
    sequential solve in every case."""
    import torch
    rng = np.random.default_rng(5200 + order)
    cases = [((136, 132, 144), [False, False, False]), ((136, 132, 144), [True, True, False]),
             ((130, 129, 160), [True, False, False]), ((136, 160), [False, False]), ((264, 136), [True, True]),
             ((136, 132, 144), [True, True, True]), ((40, 70, 160), [False, False, True]), ((40, 70, 160), [False, True, True])]
    try:
        pkg.set_sweep_path("tiled")
        for shape, per in cases:
            dim = len(shape)
            fields = 3 if dim == 2 else 1
            f = rng.standard_normal((fields,) + shape)
            f[:, 5:9] = 0.0               # exact zeros: right-hand sides the fast division path refuses
            f[:, -1] = 1e-300
            t = pkg.InterpolationFunctionTemplate(order, shape, [(0.0, 1.0 + d) for d in range(dim)], per)
            fn_dev = t.interpolate(torch.from_numpy(f if fields > 1 else f[0]).cuda())
            for k in range(fields):
                o = OracleSpline(order, shape, per, lo=[0.0] * dim, hi=[1.0 + d for d in range(dim)], f=f[k])
                assert np.array_equal(fn_dev.control_points(field=k), o.control_points()), (shape, per, k)
        f32 = rng.standard_normal((136, 132, 144)).astype(np.float32)
        fn32 = pkg.InterpolationFunction(order, f32, [(0.0, 1.0)] * 3, dtype=np.float32)
        pkg.set_sweep_path("lines")
        ref32 = pkg.InterpolationFunction(order, f32, [(0.0, 1.0)] * 3, dtype=np.float32).control_points()
        assert np.array_equal(fn32.control_points(), ref32)
    finally:
        pkg.set_sweep_path("auto")


@pytest.mark.parametrize("order,periodic", [(1, True), (2, False), (3, False), (3, True), (4, True), (5, False), (5, True)])
def test_chunk_parallel_solve_long_lines(pkg, order, periodic):
    """Few, long lines take the chunk-parallel sweep (warm-up window instead of the sequential
    dependency).  Control points must match the sequential reference algorithm to 1e-13 of
    the field scale (they are normally bit-identical: the truncation error is < 1e-23)."""
    rng = np.random.default_rng(900 + order)
    n = 200_003
    f = np.cos(np.arange(n) * 0.001) + 0.3 * rng.standard_normal(n)
    o = OracleSpline(order, (n,), [periodic], lo=[-1.0], hi=[3.0], f=f)
    fn = pkg.InterpolationFunction(order, f, [(-1.0, 3.0)], [periodic])
    c, ref = fn.control_points(), o.control_points()
    assert np.abs(c - ref).max() <= 1e-13 * np.abs(ref).max()
    assert (c != ref).mean() < 1e-3  # bit-identical except, at most, isolated last-bit ties
    pts = rng.uniform(-1.0, 3.0, 5000)
    _close(fn(pts), o.eval(pts))
    # 2-D: 3 long lines per axis-0 sweep are not enough to chunk, the long axis is
    shape = (8, 50_000)
    f2 = rng.standard_normal(shape)
    o2 = OracleSpline(order, shape, [False, periodic], lo=[0, 0], hi=[1, 1], f=f2)
    fn2 = pkg.InterpolationFunction(order, f2, [(0.0, 1.0), (0.0, 1.0)], [False, periodic])
    assert np.abs(fn2.control_points() - o2.control_points()).max() <= 1e-13 * np.abs(o2.control_points()).max()
    shape = (40_000, 8)
    f3 = rng.standard_normal(shape)
    o3 = OracleSpline(order, shape, [periodic, False], lo=[0, 0], hi=[1, 1], f=f3)
    fn3 = pkg.InterpolationFunction(order, f3, [(0.0, 1.0), (0.0, 1.0)], [periodic, False])
    assert np.abs(fn3.control_points() - o3.control_points()).max() <= 1e-13 * np.abs(o3.control_points()).max()


@pytest.mark.parametrize("dim,order,periodic", [(1, 3, (True,)), (2, 3, (False, True)), (3, 3, (False, False, False)),
                                                 (3, 2, (True, False, True)), (2, 5, (False, False))])
def test_fp32_within_1e5_of_fp64_reference(pkg, dim, order, periodic):
    """T = U = float build: values, gradient and control points within 1e-5 relative of the fp64
    oracle (north star: 1e-5 for fp32); knots are built in float arithmetic like the reference's
    coord_type = float instantiation."""
    rng = np.random.default_rng(70 + dim + order)
    shape = small_shapes(dim, order, periodic)
    lo = np.zeros(dim); hi = np.arange(1, dim + 1, dtype=np.float64)
    f = smooth_field(shape, rng).astype(np.float32)
    o = OracleSpline(order, shape, periodic, lo=lo, hi=hi, f=f.astype(np.float64))
    fn = pkg.InterpolationFunction(order, f, _ranges(lo, hi), periodic, dtype=np.float32)
    assert fn.dtype == np.float32
    ref_c = o.control_points()
    assert np.abs(fn.control_points() - ref_c).max() <= 1e-5 * np.abs(ref_c).max()
    # stay half a cell away from the ends: float knots differ from double knots by an ulp(float)
    pts = (lo + (0.02 + 0.96 * rng.uniform(0, 1, (4000, dim))) * (hi - lo)).astype(np.float32)
    ref = o.eval(pts.astype(np.float64))
    got = fn(pts)
    assert got.dtype == np.float32
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    vg = fn.value_grad(pts)
    for d in range(dim):
        dv = [0] * dim; dv[d] = 1
        g = o.deriv(pts.astype(np.float64), dv)
        assert np.abs(vg[:, 1 + d] - g).max() <= 2e-4 * np.abs(g).max()  # derivative: one order less smooth
    if dim == 3:
        import torch
        try:
            pkg.set_eval_path("binned")
            b = fn.value_grad(torch.from_numpy(pts).cuda()).cpu().numpy()
        finally:
            pkg.set_eval_path("auto")
        assert np.abs(b - vg).max() <= 1e-5 * np.abs(vg).max()


def test_many_fields_cfg5_shape(pkg):
    """cfg5 scaled down: fields on one 128x128 cubic mesh, one query set for all fields."""
    import torch
    rng = np.random.default_rng(8)
    F, shape, Q = 24, (128, 128), 5000
    fields = rng.standard_normal((F,) + shape)
    t = pkg.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0), (0.0, 1.0)])
    fn = t.interpolate(torch.from_numpy(fields).cuda())
    pts = rng.uniform(0, 1, (Q, 2))
    allv = fn.evaluate_fields(torch.from_numpy(pts).cuda()).cpu().numpy()
    for k in (0, 7, F - 1):
        o = OracleSpline(3, shape, [0, 0], lo=[0, 0], hi=[1, 1], f=fields[k])
        assert np.array_equal(fn.control_points(k), o.control_points())
        _close(allv[k], o.eval(pts))


def test_baseline_cfg1_2d_cubic_1024(pkg):
    """BASELINE configs[0]: 2-D cubic 1024x1024 mesh, 1M random queries, against the oracle."""
    rng = np.random.default_rng(12345)
    shape = (1024, 1024)
    f = smooth_field(shape, rng)
    o = OracleSpline(3, shape, [0, 0], lo=[0, 0], hi=[1, 1], f=f, nthreads=8)
    fn = pkg.InterpolationFunction(3, f, [(0.0, 1.0), (0.0, 1.0)])
    _same_control_points(fn.control_points(), o.control_points())
    pts = rng.uniform(0, 1, (1 << 20, 2))
    assert np.array_equal(fn.locate(pts), o.spans(pts))
    _close(fn(pts), o.eval(pts, 8))
    _close(fn.derivative(pts, [1, 0]), o.deriv(pts, [1, 0], 8))


def test_baseline_cfg2_1d_quintic_periodic_long(pkg):
    """BASELINE configs[1] at full size: 1-D order-5 periodic, 2^24 mesh points, value + first
    derivative; compact LU factors (bspl_host.h) and the chunk-parallel solve.  A long axis is
    where rounded knot values matter most: spans exact, tolerance still 1e-12."""
    rng = np.random.default_rng(7)
    n = 1 << 24
    f = np.sin(np.arange(n) * (14 * np.pi / n)) + 0.1 * rng.standard_normal(n)
    o = OracleSpline(5, (n,), [1], lo=[0.0], hi=[1.0], f=f)
    fn = pkg.InterpolationFunction(5, f, [(0.0, 1.0)], [True])
    assert np.array_equal(fn.knots(0), o.knots(0))
    c, ref = fn.control_points(), o.control_points()
    assert np.abs(c - ref).max() <= 1e-13 * np.abs(ref).max()
    assert (c != ref).mean() < 1e-3
    pts = rng.uniform(-0.5, 1.5, 1 << 20)  # includes wrapped queries
    assert np.array_equal(fn.locate(pts)[:, 0], o.spans(pts)[:, 0])
    vg = fn.value_grad(pts)
    _close(vg[:, 0], o.eval(pts, 8))
    _close(vg[:, 1], o.deriv(pts, [1], 8))
    idx = rng.integers(0, n, 100000)
    assert np.abs(fn(idx / n) - f[idx]).max() <= 1e-12 * np.abs(f).max()   # reproduces its data


@pytest.mark.parametrize("periodic", [False, True])
def test_baseline_cfg4_3d_cubic_512_solve(pkg, periodic):
    """BASELINE configs[3] at full size: the 512^3 cubic control-point solve (TMA-tiled contiguous
    sweep + two strided sweeps; transposing route when periodic).  Control points bit-identical to
    the oracle's sequential solve, and the spline reproduces its data at the mesh nodes."""
    import torch
    rng = np.random.default_rng(512 + periodic)
    shape = (512, 512, 512)
    f = rng.standard_normal(shape)
    per = [periodic] * 3
    o = OracleSpline(3, shape, per, lo=[0, 0, 0], hi=[1, 1, 1], f=f, nthreads=8)
    fn = pkg.InterpolationFunction(3, torch.from_numpy(f).cuda(), [(0.0, 1.0)] * 3, per)
    assert np.array_equal(fn.control_points(), o.control_points())
    idx = rng.integers(0, 512, size=(50000, 3))
    nodes = idx / (512.0 if periodic else 511.0)
    assert np.abs(fn(nodes) - f[idx[:, 0], idx[:, 1], idx[:, 2]]).max() <= 1e-12 * np.abs(f).max()
    pts = rng.uniform(0, 1, (1 << 16, 3))
    assert np.array_equal(fn.locate(pts), o.spans(pts))
    _close(fn(pts), o.eval(pts, 8))


def test_baseline_cfg3_3d_cubic_256_sample(pkg):
    """BASELINE configs[2] at full mesh size: 256^3 cubic, value + gradient through the binned/TMA
    path on 2^22 queries; a 2^17-query sample is checked against the oracle, the whole batch
    against the direct path, and the spline must reproduce its data at the mesh nodes."""
    import torch
    rng = np.random.default_rng(99)
    shape = (256, 256, 256)
    f = smooth_field(shape, rng)
    fn = pkg.InterpolationFunction(3, torch.from_numpy(f).cuda(), [(0.0, 1.0)] * 3)
    o = OracleSpline(3, shape, [0, 0, 0], lo=[0, 0, 0], hi=[1, 1, 1], f=f, nthreads=8)
    _same_control_points(fn.control_points(), o.control_points())
    pts = rng.uniform(0, 1, (1 << 22, 3))
    d = torch.from_numpy(pts).cuda()
    vg = fn.value_grad(d).cpu().numpy()           # auto -> binned (Q >= 2^20 and >= 40 * tiles)
    try:
        pkg.set_eval_path("direct")
        vd = fn.value_grad(d).cpu().numpy()
    finally:
        pkg.set_eval_path("auto")
    assert np.abs(vg - vd).max() <= 1e-13 * np.abs(vd).max()
    s = slice(0, 1 << 17)
    assert np.array_equal(fn.locate(pts[s]), o.spans(pts[s]))
    _close(vg[s, 0], o.eval(pts[s], 8))
    for k in range(3):
        dv = [0, 0, 0]; dv[k] = 1
        _close(vg[s, 1 + k], o.deriv(pts[s], dv, 8))
    idx = rng.integers(0, 256, size=(20000, 3))
    nodes = idx / 255.0
    assert np.abs(fn(nodes) - f[idx[:, 0], idx[:, 1], idx[:, 2]]).max() <= 1e-12 * np.abs(f).max()


@pytest.mark.parametrize("offset", [0.0, 10.0, 1000.0, -2.5e5])
def test_binned_path_on_offset_ranges(pkg, offset):
    """Ranges far from the origin: the reference divides by differences of rounded knot values
    (relative error eps * |t| / dx), so the unit-knot weight shortcut of interior tiles is only taken
    while |t| / dx <= 2^12 (Grid::params); beyond that the knot-based triangle reproduces the
    reference's rounding.  Value and gradient through the binned path against the oracle, 1e-12."""
    rng = np.random.default_rng(31)
    shape = (64, 64, 64)
    f = smooth_field(shape, rng)
    lo = [offset, offset - 1.0, offset + 2.0]; hi = [offset + 1.0, offset + 1.0, offset + 2.5]
    o = OracleSpline(3, shape, [0, 0, 0], lo=lo, hi=hi, f=f, nthreads=8)
    fn = pkg.InterpolationFunction(3, f, _ranges(lo, hi))
    pts = np.array(lo) + rng.uniform(0, 1, (1 << 17, 3)) * (np.array(hi) - np.array(lo))
    assert np.array_equal(fn.locate(pts), o.spans(pts))
    try:
        pkg.set_eval_path("binned")
        vg = fn.value_grad(pts)
        v = fn(pts)
    finally:
        pkg.set_eval_path("auto")
    _close(v, o.eval(pts, 8))
    _close(vg[:, 0], o.eval(pts, 8))
    for k in range(3):
        dv = [0, 0, 0]; dv[k] = 1
        _close(vg[:, 1 + k], o.deriv(pts, dv, 8))


def test_span_selection_random_ranges_bit_exact(pkg):
    """Property test (hypothesis): span - order is exactly the reference's for random orders, lengths,
    periodicities and ranges (far from the origin, tiny and huge spacings), on adversarial points --
    on knots, one ulp around them, at the range ends, far outside, several periods away."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(order=st.integers(0, 5), per=st.booleans(), n=st.integers(12, 300),
           lo=st.floats(-1e6, 1e6, allow_nan=False), width=st.floats(1e-6, 1e6, allow_nan=False), seed=st.integers(0, 1 << 30))
    def check(order, per, n, lo, width, seed):
        hi = lo + width
        if not hi > lo:
            return
        rng = np.random.default_rng(seed)
        f = rng.standard_normal(n)
        o = OracleSpline(order, (n,), [per], lo=[lo], hi=[hi], f=f)
        fn = pkg.InterpolationFunction(order, f, [(lo, hi)], [per])
        assert np.array_equal(fn.knots(0), o.knots(0))
        rlo, rhi = o.range(0)
        pts = adversarial_points([o.knots(0)], [rlo], [rhi], [per], rng, per_axis=200)
        assert np.array_equal(fn.locate(pts), o.spans(pts))
        assert np.array_equal(fn.control_points(), o.control_points())

    check()


def test_more_than_2_28_queries_are_sliced(pkg):
    """Maximum sizes: the query sort needs 32 bytes of scratch per query, so batches beyond 2^28
    queries run in slices (bspl_capi.cu: launch_eval).  2^28 + 12 345 device-resident queries; a
    sample from every slice, the tail slice included, must equal a small direct evaluation."""
    import torch
    if torch.cuda.mem_get_info()[0] < 40 * (1 << 30):
        pytest.skip("needs ~25 GB of device memory")
    rng = np.random.default_rng(2828)
    shape = (64, 64, 64)
    fn = pkg.InterpolationFunction(3, smooth_field(shape, rng), [(0.0, 1.0)] * 3)
    q = (1 << 28) + 12345
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    pts = torch.rand((q, 3), dtype=torch.float64, device="cuda", generator=gen)
    out = fn.evaluate(pts)
    idx = torch.cat([torch.randint(0, q, (200000,), device="cuda", generator=gen),
                     torch.arange(q - 12345, q, device="cuda")])
    sample = pts[idx].contiguous()
    try:
        pkg.set_eval_path("direct")
        ref = fn.evaluate(sample)
    finally:
        pkg.set_eval_path("auto")
    assert (out[idx] - ref).abs().max().item() <= 1e-13 * ref.abs().max().item()
    del pts, out
    torch.cuda.empty_cache()


def test_vector_valued_circle_as_two_fields(pkg):
    """interpolation-test.cpp:674-703 (T = Vec<2,float>, U = float): a closed curve interpolated
    componentwise; here the components are the fields of one handle sharing one query set."""
    n = 31
    theta = 2.0 * np.pi * np.arange(n) / n
    pts_xy = np.stack([np.cos(theta), np.sin(theta)]).astype(np.float32)  # [2 fields][n]
    t = pkg.InterpolationFunctionTemplate(3, (n,), [(0.0, float(n))], [True], dtype=np.float32)
    circle = t.interpolate(pts_xy)
    q = (np.arange(1024, dtype=np.float32) * (n / 1024.0)).reshape(-1, 1)  # parameter in mesh units
    xy = circle.evaluate_fields(q)
    err = np.abs(np.hypot(xy[0], xy[1]) - 1.0).mean()
    assert err < 1e-5  # the reference's bound, :697


def test_concurrent_host_threads(pkg):
    """Evaluation is re-entrant: several host threads on one handle (host-pointer path is serialised
    by the staging pipe, device results must still be exact)."""
    import threading
    rng = np.random.default_rng(4)
    shape = (20, 24, 28)
    f = smooth_field(shape, rng)
    fn = pkg.InterpolationFunction(3, f, [(0.0, 1.0)] * 3)
    pts = [rng.uniform(0, 1, (30000, 3)) for _ in range(4)]
    want = [fn.value_grad(p) for p in pts]
    got = [None] * 4

    def work(i):
        for _ in range(5):
            got[i] = fn.value_grad(pts[i])

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for w, g in zip(want, got):
        assert np.array_equal(w, g)


def test_non_finite_queries_do_not_disturb_the_batch(pkg):
    """NaN / inf coordinates must neither crash a kernel nor change the results of the other
    queries of the batch (both evaluation paths)."""
    import torch
    rng = np.random.default_rng(17)
    shape = (40, 36, 44)
    f = smooth_field(shape, rng)
    for periodic in ((False, False, False), (True, True, True)):
        fn = pkg.InterpolationFunction(3, f, [(0.0, 1.0)] * 3, periodic)
        pts = rng.uniform(0, 1, (50000, 3))
        bad = pts.copy()
        rows = rng.choice(len(pts), 500, replace=False)
        bad[rows[:200], 0] = np.nan
        bad[rows[200:350], 1] = np.inf
        bad[rows[350:], 2] = -np.inf
        good = np.ones(len(pts), dtype=bool); good[rows] = False
        for path in ("direct", "binned"):
            try:
                pkg.set_eval_path(path)
                ref = fn.value_grad(torch.from_numpy(pts).cuda()).cpu().numpy()
                got = fn.value_grad(torch.from_numpy(bad).cuda()).cpu().numpy()
            finally:
                pkg.set_eval_path("auto")
            assert np.array_equal(got[good], ref[good])
            assert np.isnan(got[rows[:200], 0]).all()


@pytest.mark.parametrize("path", ["direct", "binned"])
def test_eval_proxy_query_plan(pkg, path):
    """eval_proxy analogue: one plan (locate + tile sort), many evaluations -- other fields, other
    functions of the same template, derivatives, value+gradient, host and device outputs."""
    import torch
    rng = np.random.default_rng(23)
    shape = (36, 41, 30)
    per = [True, False, False]
    t = pkg.InterpolationFunctionTemplate(3, shape, [(0.0, 2.0), (0.0, 1.0), (-1.0, 1.0)], per)
    f1 = np.stack([smooth_field(shape, rng) for _ in range(2)])
    fa = t.interpolate(f1)
    fb = t.interpolate(smooth_field(shape, rng))
    pts = np.array([0.0, 0.0, -1.0]) + rng.uniform(0, 1, (40000, 3)) * np.array([2.0, 1.0, 2.0])
    try:
        pkg.set_eval_path(path)
        plan = fa.eval_proxy(torch.from_numpy(pts).cuda())
        plan_h = fb.eval_proxy(pts)  # built from host points
        assert np.array_equal(plan(fa, field=1), fa.evaluate(pts, field=1))
        assert np.array_equal(plan(fb), fb.evaluate(pts))
        assert np.array_equal(plan_h(fa, value_grad=True), fa.value_grad(pts))
        assert np.array_equal(plan(fb, derivatives=[1, 0, 2]), fb.derivative(pts, [1, 0, 2]))
        assert not plan(fb, derivatives=[4, 0, 0]).any()
        dev = plan(fb, value_grad=True, device_out=True)
        assert np.array_equal(dev.cpu().numpy(), fb.value_grad(pts))
        other = pkg.InterpolationFunction(3, smooth_field(shape, rng), [(0.0, 2.0), (0.0, 1.0), (-1.0, 1.0)], per)
        with pytest.raises(pkg.BsplError):
            plan(other)  # a different template
    finally:
        pkg.set_eval_path("auto")


@pytest.mark.parametrize("order,periodic", [(3, (True, False)), (2, (False, True)), (5, (False, False)), (1, (True, True))])
def test_many_fields_streamed_through_shared_memory(pkg, order, periodic):
    """evaluate_fields on a small 2-D mesh takes the field-streaming kernel (weights once per query,
    fields through shared memory): must equal per-field evaluation and the oracle."""
    import torch
    rng = np.random.default_rng(60 + order)
    F, shape, Q = 12, (48, 56), 9000
    fields = rng.standard_normal((F,) + shape)
    t = pkg.InterpolationFunctionTemplate(order, shape, [(0.0, 1.0), (-1.0, 2.0)], periodic)
    fn = t.interpolate(fields)
    pts = np.array([0.0, -1.0]) + rng.uniform(-0.1, 1.1, (Q, 2)) * np.array([1.0, 3.0])
    allv = fn.evaluate_fields(torch.from_numpy(pts).cuda()).cpu().numpy()
    host = fn.evaluate_fields(pts)
    assert np.array_equal(allv, host)
    for k in (0, 5, F - 1):
        single = fn.evaluate(pts, field=k)
        assert np.abs(allv[k] - single).max() <= 1e-13 * np.abs(single).max()
        o = OracleSpline(order, shape, periodic, lo=[0, -1], hi=[1, 2], f=fields[k])
        inside = np.all((pts >= [0, -1]) & (pts <= [1, 2]), axis=1) | np.array(periodic).all()
        _close(allv[k][inside], o.eval(pts[inside]))


@pytest.mark.parametrize("order,periodic", [(3, (False, True)), (4, (True, False))])
def test_many_fields_fp32(pkg, order, periodic):
    """The field-streaming kernel in float (32 bank classes, one per lane of a warp): equal to
    per-field float evaluation, and within the north star's 1e-5 of the float64 function."""
    rng = np.random.default_rng(80 + order)
    F, shape, Q = 9, (40, 52), 20000
    fields = rng.standard_normal((F,) + shape)
    rg = [(0.0, 1.0), (-1.0, 2.0)]
    fn32 = pkg.InterpolationFunctionTemplate(order, shape, rg, periodic, dtype=np.float32).interpolate(fields.astype(np.float32))
    fn64 = pkg.InterpolationFunctionTemplate(order, shape, rg, periodic).interpolate(fields)
    pts = np.array([0.0, -1.0]) + rng.uniform(0, 1, (Q, 2)) * np.array([1.0, 3.0])
    allv = fn32.evaluate_fields(pts.astype(np.float32))
    assert allv.dtype == np.float32 and allv.shape == (F, Q)
    ref = fn64.evaluate_fields(pts)
    for k in range(F):
        single = fn32.evaluate(pts.astype(np.float32), field=k)
        assert np.abs(allv[k] - single).max() <= 2e-6 * np.abs(single).max()
        assert np.abs(allv[k] - ref[k]).max() <= 1e-5 * np.abs(ref[k]).max() * 8  # cancellation near zeros of the field
    assert rel_err(allv, ref) <= 1e-5


def test_many_fields_host_path_is_chunked(pkg):
    """Host-pointer evaluate_fields with many fields: the staging buffers stay bounded (chunks
    shrink with the field count) and the strided copy-back lands every field in its row."""
    rng = np.random.default_rng(2)
    F, shape, Q = 600, (16, 20), 70000
    fields = rng.standard_normal((F,) + shape)
    fn = pkg.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0), (0.0, 1.0)]).interpolate(fields)
    pts = rng.uniform(0, 1, (Q, 2))
    allv = fn.evaluate_fields(pts)
    assert allv.shape == (F, Q)
    for k in (0, 299, F - 1):
        single = fn.evaluate(pts, field=k)
        assert np.abs(allv[k] - single).max() <= 1e-13 * np.abs(single).max()


# ---- many fields as a per-cell contraction (query-major results) ---------------------------------
@pytest.mark.parametrize("dim,order,periodic", [(2, 3, (False, False)), (2, 3, (True, False)), (2, 2, (False, True)),
                                                (2, 5, (True, True)), (2, 1, (False, False)), (2, 4, (False, False)),
                                                (1, 3, (False,)), (1, 5, (True,)), (3, 1, (False, True, False)),
                                                (3, 2, (True, False, False)), (3, 3, (False, False, True))])
def test_many_fields_contraction_query_major(pkg, dim, order, periodic):
    """bspl_evaluate_fields_query_major: queries sorted by cell, one (queries) x (O+1)^D x (fields) product per
    cell out of a field-minor copy of the control points.  Equal to per-field evaluation (1e-13) and to the
    oracle (1e-12); host and device entry points agree."""
    import torch
    rng = np.random.default_rng(7000 + 100 * dim + order)
    shape = {1: (61,), 2: (26, 31), 3: (9, 11, 10)}[dim]
    F, Q = 72, 30011
    fields = rng.standard_normal((F,) + shape)
    lo = [0.0, -1.0, 2.0][:dim]; hi = [1.0, 2.0, 5.0][:dim]
    fn = pkg.InterpolationFunctionTemplate(order, shape, _ranges(lo, hi), periodic).interpolate(fields)
    pts = np.array(lo) + rng.uniform(-0.05, 1.05, (Q, dim)) * (np.array(hi) - np.array(lo))
    pts[:500] = pts[0]        # one crowded cell: more queries than a work item holds
    dpts = torch.from_numpy(pts).cuda()
    try:
        pkg.set_fields_path("contract")
        qm = fn.evaluate_fields(dpts, layout="query_major").cpu().numpy()
        host = fn.evaluate_fields(pts, layout="query_major")
        fm = fn.evaluate_fields(dpts).cpu().numpy()          # field-major through the contraction + transpose
        dv = [min(order, 1)] + [0] * (dim - 1)
        dq = fn.evaluate_fields(dpts, layout="query_major", derivatives=dv).cpu().numpy()
        pkg.set_fields_path("gather")
        qm_g = fn.evaluate_fields(dpts, layout="query_major").cpu().numpy()   # field-major kernels + transpose
    finally:
        pkg.set_fields_path("auto")
    assert qm.shape == (Q, F) and fm.shape == (F, Q)
    assert np.array_equal(qm, host)
    assert np.array_equal(qm.T, fm)
    inside = np.all((pts >= lo) & (pts <= hi), axis=1) | np.array(periodic).all()
    for k in (0, 17, F - 1):
        single = fn.evaluate(pts, field=k)
        assert np.abs(qm[:, k] - single).max() <= 1e-13 * np.abs(single).max()
        assert np.abs(qm_g[:, k] - single).max() <= 1e-13 * np.abs(single).max()
        sd = fn.evaluate(pts, derivatives=dv, field=k)
        assert np.abs(dq[:, k] - sd).max() <= 1e-12 * max(np.abs(sd).max(), 1e-300)
        o = OracleSpline(order, shape, periodic, lo=lo, hi=hi, f=fields[k])
        _close(qm[inside, k], o.eval(pts[inside]))


def test_many_fields_contraction_odd_shapes_fall_back(pkg):
    """Field counts the vector accesses cannot serve (not a multiple of 4) and unaligned outputs take the
    field-major kernels plus a transpose: same numbers."""
    import torch
    rng = np.random.default_rng(91)
    F, shape, Q = 37, (20, 24), 5003
    fields = rng.standard_normal((F,) + shape)
    fn = pkg.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0), (0.0, 2.0)]).interpolate(fields)
    pts = rng.uniform(0, 1, (Q, 2)) * np.array([1.0, 2.0])
    qm = fn.evaluate_fields(torch.from_numpy(pts).cuda(), layout="query_major").cpu().numpy()
    for k in (0, 36):
        single = fn.evaluate(pts, field=k)
        assert np.abs(qm[:, k] - single).max() <= 1e-13 * np.abs(single).max()
    # re-solving into the same function refreshes the field-minor copy
    F2 = 64
    f2 = rng.standard_normal((F2,) + shape)
    t = pkg.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0), (0.0, 2.0)])
    fn2 = t.interpolate(f2)
    a = fn2.evaluate_fields(pts, layout="query_major")
    t.interpolate(2.0 * f2, into=fn2)
    b = fn2.evaluate_fields(pts, layout="query_major")
    assert np.abs(b - 2.0 * a).max() <= 1e-13 * np.abs(b).max()


def test_many_fields_contraction_fp32(pkg):
    rng = np.random.default_rng(92)
    F, shape, Q = 64, (40, 52), 20000
    fields = rng.standard_normal((F,) + shape)
    rg = [(0.0, 1.0), (-1.0, 2.0)]
    fn32 = pkg.InterpolationFunctionTemplate(3, shape, rg, dtype=np.float32).interpolate(fields.astype(np.float32))
    fn64 = pkg.InterpolationFunctionTemplate(3, shape, rg).interpolate(fields)
    pts = np.array([0.0, -1.0]) + rng.uniform(0, 1, (Q, 2)) * np.array([1.0, 3.0])
    try:
        pkg.set_fields_path("contract")
        a = fn32.evaluate_fields(pts.astype(np.float32), layout="query_major")
    finally:
        pkg.set_fields_path("auto")
    assert a.dtype == np.float32 and a.shape == (Q, F)
    assert rel_err(a, fn64.evaluate_fields(pts, layout="query_major")) <= 1e-5


# ---- the binned path's own span selection ---------------------------------------------------------
@pytest.mark.parametrize("order", [1, 2, 3, 4, 5])
@pytest.mark.parametrize("periodic", [(False, False, False), (True, False, True), (False, True, True)])
def test_binned_path_uses_the_reference_span(pkg, order, periodic):
    """The O-th derivative of a degree-O spline is constant inside a cell and jumps at every knot, so it tells
    which span a kernel used.  The cell-binned path has its own locate variants (key_of_point: locate_quick;
    eval_binned_kernel: locate_uniform_interior / locate): forced on adversarial points (on knots, +-1 ulp
    around them, range ends, wrapped by whole periods), the top derivative along each axis must be the
    oracle's -- a span off by one would differ by the size of the jump, not by rounding."""
    import torch
    rng = np.random.default_rng(8100 + order)
    shape = (33, 40, 37)
    lo, hi = axis_ranges(3, rng)
    f = rng.standard_normal(shape)          # rough data: neighbouring cells have unrelated top derivatives
    fn = pkg.InterpolationFunction(order, f, _ranges(lo, hi), periodic)
    o = OracleSpline(order, shape, periodic, lo=lo, hi=hi, f=f)
    knots = [fn.knots(d) for d in range(3)]
    rlo = np.array([o.range(d)[0] for d in range(3)]); rhi = np.array([o.range(d)[1] for d in range(3)])
    pts = np.concatenate([adversarial_points(knots, rlo, rhi, periodic, rng, per_axis=1500),
                          queries(rlo, rhi, periodic, 20000, rng, mode="wild")])
    dpts = torch.from_numpy(pts).cuda()
    spans = o.spans(pts)
    try:
        pkg.set_eval_path("binned")
        assert np.array_equal(fn.locate(pts), spans)
        for d in range(3):
            dv = [0, 0, 0]; dv[d] = order
            got = fn.evaluate(dpts, derivatives=dv).cpu().numpy()
            ref = o.deriv(pts, dv)
            jump = np.abs(np.diff(np.unique(np.round(ref, 9)))).min() if len(ref) > 1 else 1.0
            err = np.abs(got - ref)
            assert err.max() <= 1e-9 * np.abs(ref).max(), (d, err.max(), np.abs(ref).max(), jump)
        vg = fn.value_grad(dpts).cpu().numpy()
        inside = np.all((pts >= rlo) & (pts <= rhi), axis=1) | np.array(periodic).all()
        _close(vg[inside, 0], o.eval(pts[inside]))
    finally:
        pkg.set_eval_path("auto")


def test_plan_value_grad_rejects_unaligned_device_output(pkg):
    """A tiled plan stores {value, gradient} as one 32-byte vector: an unaligned device `out` is refused
    instead of faulting."""
    import torch
    rng = np.random.default_rng(5)
    shape = (40, 40, 40)
    fn = pkg.InterpolationFunction(3, rng.standard_normal(shape), [(0.0, 1.0)] * 3)
    pts = torch.rand((1 << 16, 3), dtype=torch.float64, device="cuda")
    try:
        pkg.set_eval_path("binned")
        plan = fn.eval_proxy(pts)
        buf = torch.empty((1 << 16) * 4 + 1, dtype=torch.float64, device="cuda")
        ok = plan(fn, value_grad=True, out=buf[:-1].view(-1, 4))
        with pytest.raises(pkg.BsplError):
            plan(fn, value_grad=True, out=buf[1:].view(-1, 4))
        assert torch.isfinite(ok).all()
    finally:
        pkg.set_eval_path("auto")


def test_wrong_length_coordinate_arrays_are_refused(pkg):
    with pytest.raises(ValueError):
        pkg.InterpolationFunctionTemplate(3, (20, 24), [np.linspace(0, 1, 19), (0.0, 1.0)])
    # a two-element list on an axis that is not two points long is a (min, max) range
    t = pkg.InterpolationFunctionTemplate(3, (20, 24), [[0.0, 1.0], [0.0, 2.0]])
    assert t.shape == (20, 24)
