"""CPU-only: the C-ABI library builds for sm_100a, loads, exports every symbol
include/bspline_b200.h declares, and its host-side template construction (knots,
collocation LU) equals the oracle bit for bit.  No kernels run here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle.pyoracle import OracleSpline

dp = C.POINTER(C.c_double)


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bspline_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bspl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib_built):
    from bsplineinterpolation_b200 import _capi
    names = _declared_symbols()
    assert len(names) >= 20
    L = lib_built.lib()
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert set(names) == set(_capi.SIGNATURES), "python signatures out of sync with the header"
    assert b"sm_100a" in L.bspl_version()


def test_header_is_plain_c(tmp_path, lib_built):
    """include/bspline_b200.h is a C header (C99, -pedantic): what cgo / JNI / ctypes-style bindings need."""
    import subprocess
    src = tmp_path / "c_abi.c"
    src.write_text("#include <bspline_b200.h>\n"
                   "int main(void) { bspl_template* t = 0; bspl_function* f = 0; bspl_query_plan* p = 0;\n"
                   "  (void)t; (void)f; (void)p; return bspl_version() == 0 || BSPL_OK != 0; }\n")
    exe = tmp_path / "c_abi"
    pkg = os.path.join(ROOT, "bsplineinterpolation_b200")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", pkg, "-lbspline_b200", "-Wl,-rpath," + pkg])
    assert subprocess.run([str(exe)]).returncode == 0


def test_no_cpu_fallback_in_product():
    """The product package must not import or link the oracle."""
    pkg = os.path.join(ROOT, "bsplineinterpolation_b200")
    for dirpath, _d, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.replace("no oracle", ""), fn


def _host_knots(L, order, per, n, lo, hi, coords=None):
    from bsplineinterpolation_b200._capi import check
    nk = C.c_int64(); rng = (C.c_double * 2)()
    cp = coords.ctypes.data_as(dp) if coords is not None else None
    check(L.bspl_host_axis_knots(0, order, per, n, lo, hi, cp, None, 0, C.byref(nk), rng))
    out = np.empty(nk.value)
    check(L.bspl_host_axis_knots(0, order, per, n, lo, hi, cp, out.ctypes.data_as(dp), out.size, C.byref(nk), rng))
    return out, (rng[0], rng[1])


def _host_factor(L, order, per, n, lo, hi, coords=None):
    from bsplineinterpolation_b200._capi import check
    band = C.c_int()
    cp = coords.ctypes.data_as(dp) if coords is not None else None
    check(L.bspl_host_axis_factor(0, order, per, n, lo, hi, cp, C.byref(band), None, None, None, None, None))
    P = band.value
    arrs = [np.zeros((n, max(P, 1))) for _ in range(2)] + [np.zeros(n)] + [np.zeros((n, max(P, 1))) for _ in range(2)]
    check(L.bspl_host_axis_factor(0, order, per, n, lo, hi, cp, C.byref(band), *[a.ctypes.data_as(dp) for a in arrs]))
    return (P,) + tuple(arrs)


def _row_form_from_oracle(o, n, P, per):
    """The oracle's diagonal-major LU store (BandMatrix.hpp:50-60, :122-146) in the row form the
    kernels consume."""
    band, right, bottom = o.lu(0)
    p = band.shape[1] // 2
    Lr = np.zeros((n, max(P, 1))); U = np.zeros_like(Lr)
    dg = band[:, p].copy()
    i = np.arange(n)
    for m in range(P):
        jl = i - P + m
        ok = jl >= 0
        Lr[ok, m] = band[jl[ok], (i + p - jl)[ok]]
        ju = i + 1 + m
        ok = ju < n
        U[ok, m] = band[ju[ok], (i + p - ju)[ok]]
    B = np.zeros_like(Lr); R = np.zeros_like(Lr)
    if per and P > 0:
        for c in range(p):            # right strip: column j = n-p+c, rows i < j-p
            j = n - p + c
            rows = np.arange(0, j - p)
            R[rows, j - (n - P)] = right[rows, j + p - n]
        for r in range(p):            # bottom strip: row ii = n-p+r, columns j < ii-p
            ii = n - p + r
            cols = np.arange(0, ii - p)
            B[cols, ii - (n - P)] = bottom[cols, ii + p - n]
    return Lr, U, dg, B, R


@pytest.mark.parametrize("order", range(6))
@pytest.mark.parametrize("per", [0, 1])
def test_host_template_construction_matches_oracle(lib_built, order, per):
    """Knots, ranges and LU factors bit-identical to the oracle's -- including long uniform axes,
    where the host factorisation fast-forwards through its steady state (n = 1500, 4099) or keeps only
    both ends of the matrix (compact form, n = 20011)."""
    L = lib_built.lib()
    rng = np.random.default_rng(order * 2 + per)
    for n in (7, 12, 33, 64, 1500, 4099, 16383, 16384, 20011):
        for nonuni in (0, 1):
            if nonuni and ((order == 0 and not per) or n > 64):
                continue
            lo, hi = -1.3, 2.9
            coords = None
            if nonuni:
                coords = np.sort(rng.uniform(lo, hi, n + per)); coords[0] = lo; coords[-1] = hi
            o = OracleSpline(order, (n,), [per], lo=[lo], hi=[hi], coords=[coords] if nonuni else None)
            k, r = _host_knots(L, order, per, n, lo, hi, coords)
            assert np.array_equal(k, o.knots(0)) and r == o.range(0)
            P, Lr, U, dg, B, R = _host_factor(L, order, per, n, lo, hi, coords)
            eL, eU, edg, eB, eR = _row_form_from_oracle(o, n, P, per)
            assert P == o.lu(0)[0].shape[1] // 2
            for got, exp, name in ((Lr, eL, "L"), (U, eU, "U"), (dg, edg, "diag"), (B, eB, "bottom"), (R, eR, "right")):
                assert np.array_equal(got, exp), (name, order, per, n, nonuni)


def test_invalid_arguments_reported(lib_built):
    from bsplineinterpolation_b200 import BsplError, InterpolationFunctionTemplate
    with pytest.raises(BsplError) as e:
        InterpolationFunctionTemplate(9, (8,), [(0.0, 1.0)])
    assert e.value.code == 5  # BSPL_ERR_UNSUPPORTED
    with pytest.raises(BsplError):
        InterpolationFunctionTemplate(3, (8, 8, 8, 8), [(0.0, 1.0)] * 4)


def test_host_construction_random_parameters(lib_built):
    """Property test: for random (order, periodicity, length, range) the host-side knots, range and LU
    factors are bit-identical to the oracle's -- including ranges far from the origin and tiny or huge
    spacings, where every rounding of `a + (i - extra/2) * dx` matters."""
    from hypothesis import given, settings, strategies as st

    L = lib_built.lib()

    @settings(max_examples=80, deadline=None, derandomize=True)
    @given(order=st.integers(0, 5), per=st.integers(0, 1), n=st.integers(12, 400),
           lo=st.floats(-1e6, 1e6, allow_nan=False), width=st.floats(1e-6, 1e6, allow_nan=False))
    def check(order, per, n, lo, width):
        hi = lo + width
        if not hi > lo:
            return
        o = OracleSpline(order, (n,), [per], lo=[lo], hi=[hi])
        k, r = _host_knots(L, order, per, n, lo, hi)
        assert np.array_equal(k, o.knots(0)) and r == o.range(0)
        P, Lr, U, dg, B, R = _host_factor(L, order, per, n, lo, hi)
        eL, eU, edg, eB, eR = _row_form_from_oracle(o, n, P, per)
        for got, exp in ((Lr, eL), (U, eU), (dg, edg), (B, eB), (R, eR)):
            assert np.array_equal(got, exp)

    check()
