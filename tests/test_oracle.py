"""CPU-only: pin the oracle (C restatement) against the reference's own golden
vectors and against the unmodified reference headers compiled into oracle/_ref."""
import os

import numpy as np
import pytest

from conftest import ROOT, rel_err
from cases import adversarial_points, all_combos, axis_ranges, extended_combos, queries, small_shapes, smooth_field
from oracle import pyoracle
from oracle.pyoracle import OracleSpline

TOL = 1e-14  # interpolation-test.cpp:16
G = None


@pytest.fixture(scope="module", autouse=True)
def _build():
    pyoracle.build()


def test_golden_1d(golden):
    g = golden["interpolation"]
    f = np.array(g["f"])
    cubic = OracleSpline(3, (13,), [0], lo=[0], hi=[6.0], f=f)  # x range (0, .5*(size-1)), :28
    assert rel_err(cubic.eval(g["coords_1d_half"]), g["vals_1d"]) < TOL
    assert rel_err(cubic.deriv(g["coords_1d_half"], [1]), g["vals_1d_derivative_1"]) < TOL
    for x, v in (g["extrapolate_left"], g["extrapolate_right"]):  # :92-101
        assert abs(cubic.eval([x])[0] - v) < TOL
    xs = np.array(g["coords_1d"])
    lin = OracleSpline(1, (13,), [0], lo=[0], hi=[12.0], f=f)
    idx = np.floor(xs).astype(int)
    assert rel_err(lin.eval(xs), f[idx] + (f[idx + 1] - f[idx]) * (xs - idx)) < TOL
    seg = OracleSpline(0, (13,), [0], lo=[0], hi=[12.0], f=f)
    assert rel_err(seg.eval(xs), f[np.round(xs).astype(int)]) < TOL


def test_golden_periodic_1d(golden):
    g = golden["interpolation"]
    f = np.array(g["f"])[:-1]
    xs = np.array(g["coords_1d"])
    q = OracleSpline(4, (12,), [1], lo=[0], hi=[12.0], f=f)
    assert rel_err(q.eval(xs), g["vals_1d_periodic"]) < TOL
    assert rel_err(q.deriv(xs, [1]), g["vals_1d_derivative_periodic"]) < TOL
    assert abs(q.eval([xs[0] - 12])[0] - g["vals_1d_periodic"][0]) < TOL  # :379-384
    assert abs(q.eval([xs[0] + 12])[0] - g["vals_1d_periodic"][0]) < TOL
    idx = np.floor(xs).astype(int)
    lin = OracleSpline(1, (12,), [1], lo=[0], hi=[12.0], f=f)
    assert rel_err(lin.eval(xs), f[idx] + (f[(idx + 1) % 12] - f[idx]) * (xs - idx)) < TOL


def test_golden_2d_3d(golden):
    g = golden["interpolation"]
    f2 = np.array(g["f2"]).reshape(5, 5)
    c2 = np.array(g["coords_2d"]).reshape(-1, 2)
    s = OracleSpline(3, (5, 5), [0, 0], lo=[0, 0], hi=[4.0, 4.0], f=f2)
    assert rel_err(s.eval(c2), g["vals_2d"]) < TOL
    assert rel_err(s.deriv(c2, [2, 1]), g["vals_2d_derivative_x2_y1"]) < TOL
    sp = OracleSpline(3, (5, 4), [0, 1], lo=[0, 0], hi=[4.0, 4.0], f=f2[:, :4])
    assert rel_err(sp.eval(c2), g["vals_2d_periodic"]) < TOL
    f3 = np.array(g["f3"]).reshape(5, 6, 7)
    c3 = np.array(g["coords_3d"]).reshape(-1, 3)
    s3 = OracleSpline(3, (5, 6, 7), [0, 0, 0], lo=[0, 0, 0], hi=[4.0, 5.0, 6.0], f=f3)
    assert rel_err(s3.eval(c3), g["vals_3d"]) < TOL
    assert rel_err(s3.deriv(c3, [1, 0, 3]), g["vals_3d_derivative_x1_y0_z3"]) < TOL


def test_golden_nonuniform(golden):
    g = golden["interpolation"]
    f = np.array(g["f"])
    xc = np.array(g["input_coords_1d"])
    xs = np.array(g["coords_1d"])
    s = OracleSpline(3, (13,), [0], coords=[xc], f=f)
    assert rel_err(s.eval(xs), g["vals_1d_nonuniform"]) < TOL
    sp = OracleSpline(4, (12,), [1], coords=[xc], f=f[:-1])
    assert rel_err(sp.eval(xs), g["vals_1d_nonuniform_periodic"]) < TOL
    f2 = np.array(g["f2"]).reshape(5, 5)
    c2 = np.array(g["coords_2d"]).reshape(-1, 2)
    s2 = OracleSpline(3, (4, 5), [1, 0], lo=[0, 0], hi=[4.0, 4.0],
                      coords=[None, np.array(g["nonuniform_coord_for_2d"])], f=f2[:4])
    assert rel_err(s2.eval(c2), g["vals_2d_X_periodic_Y_nonuniform"]) < TOL


def test_golden_bspline(golden):
    """bspline-test.cpp: splines straight from knots + control points, tol 1e-15 (:45)."""
    b = golden["bspline"]
    tol = 1e-15
    k = b["knots"]
    s1 = OracleSpline.from_knots(3, [0], [k], np.array(b["cp"]))
    x1 = np.array(b["coords_1d"])
    assert rel_err(s1.eval(x1), b["vals_1d"]) < tol
    assert rel_err(s1.deriv(x1, [0]), b["vals_1d"]) < tol
    assert rel_err(s1.deriv(x1, [1]), b["vals_1d_derivative_1"]) < tol
    assert rel_err(s1.deriv(x1, [2]), b["vals_1d_derivative_2"]) < tol
    cp2 = np.array(b["cp2"]).reshape(5, 5)
    x2 = np.array(b["coords_2d"]).reshape(-1, 2)
    s2 = OracleSpline.from_knots(3, [0, 0], [k, k], cp2)
    assert rel_err(s2.eval(x2), b["vals_2d"]) < tol
    assert rel_err(s2.deriv(x2, [2, 0]), b["vals_2d_derivative_x2_y0"]) < tol
    assert rel_err(s2.deriv(x2, [1, 1]), b["vals_2d_derivative_x1_y1"]) < tol
    s2p = OracleSpline.from_knots(3, [0, 1], [k, b["knots2"]], cp2)
    assert rel_err(s2p.eval(x2), b["vals_2d_periodic"]) < tol
    assert rel_err(s2p.deriv(x2, [1, 1]), b["vals_2d_periodic_derivative_x1_y1"]) < tol
    cp3 = np.array(b["cp3"]).reshape(5, 5, 5)
    x3 = np.array(b["coords_3d"]).reshape(-1, 3)
    s3 = OracleSpline.from_knots(3, [0, 0, 0], [k, k, k], cp3)
    assert rel_err(s3.eval(x3), b["vals_3d"]) < tol


def _band_matrices(n):
    """The four matrices of band-matrix-and-solver-test.cpp:55-106."""
    lap = np.zeros((n, n))
    b2 = np.zeros((n, n)); b4 = np.zeros((n, n)); asym = np.zeros((n, n))
    for i in range(n):
        lap[i, i] = -2
        if i > 0: lap[i, i - 1] = 1
        if i < n - 1: lap[i, i + 1] = 1
        b2[i, i] = 3 / 4; b2[i, (i - 1) % n] = 1 / 8; b2[i, (i + 1) % n] = 1 / 8
        b4[i, i] = 115 / 192
        b4[i, (i - 1) % n] = b4[i, (i + 1) % n] = 19 / 96
        b4[i, (i - 2) % n] = b4[i, (i + 2) % n] = 1 / 384
        asym[i, (i - 1) % n] = 2889 / 16000; asym[i, i] = 1701 / 3200; asym[i, (i + 1) % n] = 33 / 128
        asym[i, (i + 2) % n] = 729 / 20000; asym[i, (i + 3) % n] = 9 / 20000
    return [(lap, 1, 1, 0), (b2, 1, 1, 1), (b4, 2, 2, 1), (asym, 1, 3, 1)]


def test_band_solver_residual():
    n = 64
    rhs = np.random.default_rng(5).uniform(-1, 1, n)
    for a, p, q, cyc in _band_matrices(n):
        x = pyoracle.port_band_solve(a, rhs, p, q, cyc)
        assert np.linalg.norm(a @ x - rhs) / np.linalg.norm(rhs) < 1e-10  # :30
        if pyoracle.ref_available():
            assert np.array_equal(x, pyoracle.ref_band_solve(a, rhs, p, q, cyc))


needs_ref = pytest.mark.skipif(not pyoracle.ref_available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("dim,order,periodic", list(all_combos()) + sorted(set(extended_combos())))
def test_port_equals_reference(dim, order, periodic):
    """knots, ranges, spans and control points bit-identical; values/derivatives <= 1e-13."""
    from oracle.pyoracle import RefSpline
    rng = np.random.default_rng(1000 * dim + 10 * order + sum(periodic))
    shape = small_shapes(dim, order, periodic)
    lo, hi = axis_ranges(dim, rng)
    f = smooth_field(shape, rng)
    o = OracleSpline(order, shape, periodic, lo=lo, hi=hi, f=f)
    r = RefSpline(order, f, periodic, lo=lo, hi=hi, kind="cell")
    rp = RefSpline(order, f, periodic, lo=lo, hi=hi, kind="plain")
    for d in range(dim):
        assert np.array_equal(o.knots(d), r.knots(d))
        assert o.range(d) == tuple(r.range(d))
    assert np.array_equal(o.control_points(), rp.control_points())
    knots = [o.knots(d) for d in range(dim)]
    rlo = [o.range(d)[0] for d in range(dim)]
    rhi = [o.range(d)[1] for d in range(dim)]
    adv = adversarial_points(knots, rlo, rhi, periodic, rng)
    assert np.array_equal(o.spans(adv), r.spans(adv))
    pts = queries(np.array(rlo), np.array(rhi), periodic, 500, rng, mode="wild")
    assert np.array_equal(o.spans(pts), r.spans(pts))
    assert np.array_equal(o.eval(pts), r.eval(pts))
    for dv in ([1] + [0] * (dim - 1), [min(order, 2)] * dim, [order + 1] + [0] * (dim - 1)):
        assert np.array_equal(o.deriv(pts, dv), r.deriv(pts, dv))


@needs_ref
def test_port_equals_reference_nonuniform():
    from oracle.pyoracle import RefSpline
    rng = np.random.default_rng(77)
    for order in range(1, 6):
        for per in (False, True):
            n = 29
            xc = np.sort(rng.uniform(0, 5, n + per)); xc[0] = 0; xc[-1] = 5
            f = rng.standard_normal(n)
            o = OracleSpline(order, (n,), [per], coords=[xc], f=f)
            r = RefSpline(order, f, [per], coords=[xc], kind="plain")
            assert np.array_equal(o.knots(0), r.knots(0))
            assert np.array_equal(o.control_points(), r.control_points())
            pts = rng.uniform(0, 5, 300)
            assert np.array_equal(o.eval(pts), r.eval(pts))


def test_committed_reference_outputs():
    """Outputs of the reference itself, generated here by tests/golden/make_ref_outputs.py,
    replayed against the port (works without oracle/_ref and without /root/reference)."""
    path = os.path.join(ROOT, "tests", "golden", "ref_outputs.npz")
    data = np.load(path)
    n_cases = int(data["n_cases"])
    assert n_cases >= 20
    for c in range(n_cases):
        order = int(data["c%d_order" % c]); per = [bool(v) for v in data["c%d_periodic" % c]]
        f = data["c%d_f" % c]; lo = data["c%d_lo" % c]; hi = data["c%d_hi" % c]
        o = OracleSpline(order, f.shape, per, lo=lo, hi=hi, f=f)
        assert np.array_equal(o.control_points(), data["c%d_ctrl" % c])
        pts = data["c%d_pts" % c]
        assert np.array_equal(o.spans(pts), data["c%d_spans" % c])
        assert np.array_equal(o.eval(pts), data["c%d_vals" % c])
        dv = [int(v) for v in data["c%d_dv" % c]]
        assert np.array_equal(o.deriv(pts, dv), data["c%d_dvals" % c])


def test_long_uniform_axis_matches_committed_reference_outputs():
    """tests/golden/ref_outputs_long.npz: the unmodified reference on a 16 411-point axis."""
    from cases import long_axis_field
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_outputs_long.npz"))
    n = int(d["n"])
    f = long_axis_field(n)
    for c in range(int(d["n_cases"])):
        order, per = int(d["c%d_order" % c]), bool(d["c%d_periodic" % c])
        o = OracleSpline(order, (n,), [per], lo=[-1.5], hi=[2.25], f=f)
        assert np.array_equal(o.control_points(), d["c%d_ctrl" % c])
        pts = d["c%d_pts" % c]
        assert np.array_equal(o.spans(pts).reshape(-1), d["c%d_spans" % c].reshape(-1))
        ref = d["c%d_vals" % c]
        assert np.abs(o.eval(pts) - ref).max() <= 1e-13 * np.abs(ref).max()
