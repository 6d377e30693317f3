// Drop-in check of include/intp_b200/Interpolation.hpp: the scenarios of the reference's
// test/src/interpolation-test.cpp, written against the same class names, with the
// Mathematica golden vectors supplied through golden_vectors.inc (generated from
// tests/golden/reference_vectors.json by tests/test_cpp_dropin.py).  Needs a GPU.
#define INTP_PERIODIC_NO_DUMMY_POINT  // the reference's own test configuration (test/CMakeLists.txt:46)
#include <intp_b200/Interpolation.hpp>

#include <cmath>
#include <cstdio>
#include <vector>

#include "golden_vectors.inc"  // namespace golden { const std::vector<double> f, coords_1d, ...; }

using namespace intp;

static int failures = 0;
static void expect(bool ok, const char* what, double err = 0) {
    std::printf("%-58s %s  (%.3g)\n", what, ok ? "ok" : "FAILED", err);
    if (!ok) ++failures;
}
template <typename F>
static double rel_err(F&& f, const std::vector<double>& pts, std::size_t dim, const std::vector<double>& vals) {
    double e = 0, l2 = 0;
    for (std::size_t i = 0; i < vals.size(); ++i) {
        const double v = f(&pts[i * dim]);
        e += (v - vals[i]) * (v - vals[i]);
        l2 += vals[i] * vals[i];
    }
    return std::sqrt(e / l2);
}

int main() {
    constexpr double tol = 1e-14;
    using namespace golden;

    // 1-D, orders 3 / 1 / 0, extrapolation
    InterpolationFunction1D<3> interp1{std::make_pair(0, .5 * (f.size() - 1)), util::get_range(f)};
    double d = rel_err([&](const double* x) { return interp1(x[0]); }, coords_1d_half, 1, vals_1d);
    expect(d < tol, "1D cubic", d);
    expect(std::abs(interp1(-.5) - extrapolate_left[1]) < tol, "1D extrapolation left");
    expect(std::abs(interp1(6.5) - extrapolate_right[1]) < tol, "1D extrapolation right");
    d = rel_err([&](const double* x) { return interp1.derivative_at(std::make_pair(x[0], 1)); }, coords_1d_half, 1,
                vals_1d_derivative_1);
    expect(d < tol, "1D derivative", d);
    InterpolationFunction1D<1> lin(util::get_range(f));
    d = rel_err([&](const double* x) { return lin(x[0]); }, coords_1d, 1, [&] {
        std::vector<double> v;
        for (double c : coords_1d) { auto i = static_cast<std::size_t>(std::floor(c)); v.push_back(f[i] + (f[i + 1] - f[i]) * (c - double(i))); }
        return v; }());
    expect(d < tol, "1D linear", d);

    // 1-D periodic quartic (INTP_PERIODIC_NO_DUMMY_POINT semantics)
    std::vector<double> fp(f.begin(), f.end() - 1);
    InterpolationFunction1D<4> interp1p(util::get_range(fp), true);
    d = rel_err([&](const double* x) { return interp1p(x[0]); }, coords_1d, 1, vals_1d_periodic);
    expect(d < tol, "1D periodic quartic", d);
    expect(std::abs(interp1p(coords_1d[0] - 12.) - vals_1d_periodic[0]) < tol, "periodic wrap left");
    expect(std::abs(interp1p(coords_1d[0] + 12.) - vals_1d_periodic[0]) < tol, "periodic wrap right");
    d = rel_err([&](const double* x) { return interp1p.derivative_at(std::make_pair(x[0], 1)); }, coords_1d, 1,
                vals_1d_derivative_periodic);
    expect(d < tol, "1D periodic derivative", d);

    // 2-D 5x5 cubic, bounds check, periodic y
    Mesh<double, 2> f2d{5, 5};
    for (std::size_t i = 0; i < 5; ++i) for (std::size_t j = 0; j < 5; ++j) f2d(i, j) = f2[i * 5 + j];
    InterpolationFunction<double, 2, 3> interp2{f2d, std::make_pair(0., 4.), std::make_pair(0., 4.)};
    expect(interp2.uniform(0) && interp2.uniform(1), "uniform flags");
    d = rel_err([&](const double* x) { return interp2(x[0], x[1]); }, coords_2d, 2, vals_2d);
    expect(d < tol, "2D cubic", d);
    d = rel_err([&](const double* x) { return interp2.derivative_at(std::array<double, 2>{x[0], x[1]}, 2, 1); }, coords_2d,
                2, vals_2d_derivative_x2_y1);
    expect(d < tol, "2D derivative (2,1)", d);
    bool threw = false;
    try { interp2.at(-1, 1); } catch (const std::domain_error&) { threw = true; }
    expect(threw, "at() throws std::domain_error out of range");
    Mesh<double, 2> f2py{5, 4};
    for (std::size_t i = 0; i < 5; ++i) for (std::size_t j = 0; j < 4; ++j) f2py(i, j) = f2[i * 5 + j];
    InterpolationFunction<double, 2, 3> interp2p({false, true}, f2py, std::make_pair(0., 4.), std::make_pair(0., 4.));
    d = rel_err([&](const double* x) { return interp2p(x[0], x[1]); }, coords_2d, 2, vals_2d_periodic);
    expect(d < tol, "2D periodic y", d);

    // 3-D 5x6x7 cubic + derivative (1,0,3), template reuse, batched entry points
    Mesh<double, 3> f3d{5, 6, 7};
    for (std::size_t i = 0; i < 210; ++i) f3d(f3d.dimension().dimwise_indices(i)) = f3[i];
    InterpolationFunctionTemplate<double, 3, 3> tmpl(f3d.dimension(), std::make_pair(0., 4.), std::make_pair(0., 5.),
                                                     std::make_pair(0., 6.));
    auto interp3 = tmpl.interpolate(f3d);
    d = rel_err([&](const double* x) { return interp3(x[0], x[1], x[2]); }, coords_3d, 3, vals_3d);
    expect(d < tol, "3D cubic through a template", d);
    d = rel_err([&](const double* x) { return interp3.derivative_at(std::array<double, 3>{x[0], x[1], x[2]}, {1, 0, 3}); },
                coords_3d, 3, vals_3d_derivative_x1_y0_z3);
    expect(d < tol, "3D derivative (1,0,3)", d);
    std::vector<double> batch(vals_3d.size());
    interp3.evaluate(coords_3d.data(), batch.size(), batch.data());
    double e = 0, l2 = 0;
    for (std::size_t i = 0; i < batch.size(); ++i) { e += (batch[i] - vals_3d[i]) * (batch[i] - vals_3d[i]); l2 += vals_3d[i] * vals_3d[i]; }
    expect(std::sqrt(e / l2) < tol, "batched evaluate(points, out)", std::sqrt(e / l2));
    std::vector<double> vg(4 * vals_3d.size());
    interp3.evaluate_value_grad(coords_3d.data(), vals_3d.size(), vg.data());
    double worst = 0;
    for (std::size_t i = 0; i < vals_3d.size(); ++i) {
        std::array<double, 3> c{coords_3d[3 * i], coords_3d[3 * i + 1], coords_3d[3 * i + 2]};
        worst = std::max(worst, std::abs(vg[4 * i] - interp3(c)));
        worst = std::max(worst, std::abs(vg[4 * i + 1] - interp3.derivative(c, 1, 0, 0)));
        worst = std::max(worst, std::abs(vg[4 * i + 2] - interp3.derivative(c, 0, 1, 0)));
        worst = std::max(worst, std::abs(vg[4 * i + 3] - interp3.derivative(c, 0, 0, 1)));
    }
    expect(worst < 1e-13, "fused value+gradient == separate calls", worst);
    auto copy = interp3;  // value semantics
    expect(copy(coords_3d[0], coords_3d[1], coords_3d[2]) == interp3(coords_3d[0], coords_3d[1], coords_3d[2]), "copy");
    {   // eval_proxy: weights/location once, applied to two functions of the same template
        Mesh<double, 3> g3d{5, 6, 7};
        for (std::size_t i = 0; i < 210; ++i) g3d(g3d.dimension().dimwise_indices(i)) = 2. * f3[i] - 1.;
        auto other = tmpl.interpolate(g3d);
        std::array<double, 3> c{coords_3d[0], coords_3d[1], coords_3d[2]};
        auto proxy = interp3.eval_proxy(c);
        expect(proxy(interp3) == interp3(c) && proxy(other) == other(c), "eval_proxy (single point, two functions)");
        auto batch_proxy = interp3.eval_proxy(coords_3d.data(), vals_3d.size());
        std::vector<double> pv(vals_3d.size());
        batch_proxy(other, pv.data());
        bool same = true;
        for (std::size_t i = 0; i < pv.size(); ++i)
            same = same && pv[i] == other(coords_3d[3 * i], coords_3d[3 * i + 1], coords_3d[3 * i + 2]);
        expect(same, "eval_proxy (batched)");
    }
    {   // template-level eval_proxy (InterpolationTemplate.hpp:145-165): made before any field exists
        std::array<double, 3> c{coords_3d[3], coords_3d[4], coords_3d[5]};
        decltype(tmpl)::eval_proxy_t early = tmpl.eval_proxy(c);
        auto early2 = tmpl.eval_proxy(c[0], c[1], c[2]);
        auto later = tmpl.interpolate(f3d);
        expect(early(later) == later(c) && early2(interp3) == interp3(c), "template eval_proxy before interpolate");
    }
    {   // spline(): a BSpline view on the function's own device storage
        const auto& sp = interp3.spline();
        std::array<double, 3> c{coords_3d[0], coords_3d[1], coords_3d[2]};
        bool ok = sp(c) == interp3(c) && sp.periodicity(0) == interp3.periodicity(0) && sp.get_order() == 3 &&
                  sp.range(2) == interp3.range(2) && sp.knots_num(1) == 6 + 4;
        ok = ok && sp.derivative_at({std::make_pair(c[0], std::size_t{1}), std::make_pair(c[1], std::size_t{0}),
                                     std::make_pair(c[2], std::size_t{3})}) == interp3.derivative(c, 1, 0, 3);
        const auto ctrl = interp3.control_points();
        ok = ok && sp.control_points().size() == 210 && sp.control_points()(1, 2, 3) == ctrl(1, 2, 3);
        // the same spline rebuilt from its knots and control points (BSpline.hpp:188-210)
        BSpline<double, 3, 3> rebuilt(sp.control_points(), std::make_pair(sp.knots_begin(0), sp.knots_end(0)),
                                      std::make_pair(sp.knots_begin(1), sp.knots_end(1)),
                                      std::make_pair(sp.knots_begin(2), sp.knots_end(2)));
        ok = ok && std::abs(rebuilt(c) - sp(c)) <= 1e-14 * std::abs(sp(c));  // generic-knot vs uniform-axis kernels
        auto ev = rebuilt.pre_calc_coef({std::make_pair(c[0], std::size_t{3}), std::make_pair(c[1], std::size_t{3}),
                                         std::make_pair(c[2], std::size_t{3})});
        ok = ok && ev(rebuilt) == rebuilt(c);
        expect(ok, "spline() view, BSpline from knots, pre_calc_coef");
    }
    {   // Mesh line iterators (Mesh.hpp:342-371) feeding a 1-D interpolation: column 2 of f2-like data
        Mesh<double, 2> m2{5, 5};
        for (std::size_t i = 0; i < 25; ++i) m2(m2.dimension().dimwise_indices(i)) = std::sin(0.3 * double(i));
        InterpolationFunction<double, 1, 3> col(std::make_pair(m2.begin(0, {0, 2}), m2.end(0, {0, 2})), std::make_pair(0., 4.));
        expect(std::abs(col(3.) - m2(3, 2)) < 1e-14, "1-D interpolation over a Mesh line iterator");
    }
    InterpolationFunction<double, 3, 3> into;
    tmpl.interpolate(into, f3d);
    expect(into(1., 2., 3.) == interp3(1., 2., 3.), "interpolate(function&, mesh)");
    expect(!interp3.periodicity(0) && !interp3.periodicity(1) && !interp3.periodicity(2), "periodicity()");
    expect(interp3.range(1).first == 0. && interp3.range(1).second == 5., "range()");

    // non-uniform axes
    InterpolationFunction1D<3> nu(util::get_range(input_coords_1d), util::get_range(f));
    d = rel_err([&](const double* x) { return nu(x[0]); }, coords_1d, 1, vals_1d_nonuniform);
    expect(d < tol && !nu.uniform(0), "1D non-uniform cubic", d);
    InterpolationFunction1D<4> nup(util::get_range(input_coords_1d), util::get_range(fp), true);
    d = rel_err([&](const double* x) { return nup(x[0]); }, coords_1d, 1, vals_1d_nonuniform_periodic);
    expect(d < tol, "1D non-uniform periodic quartic", d);
    Mesh<double, 2> f2px{4, 5};
    for (std::size_t i = 0; i < 4; ++i) for (std::size_t j = 0; j < 5; ++j) f2px(i, j) = f2[i * 5 + j];
    InterpolationFunction<double, 2, 3> mixed({true, false}, f2px, std::make_pair(0., 4.),
                                              util::get_range(nonuniform_coord_for_2d));
    d = rel_err([&](const double* x) { return mixed(x[0], x[1]); }, coords_2d, 2, vals_2d_X_periodic_Y_nonuniform);
    expect(d < tol, "2D x-periodic, y non-uniform", d);

    // float build (T = U = float): circle, interpolation-test.cpp:674-703 reduced to one component
    {
        constexpr std::size_t n = 31;
        std::vector<float> cx;
        for (std::size_t i = 0; i < n; ++i) cx.push_back(std::cos(2.f * 3.14159265358979f * float(i) / n));
        InterpolationFunction1D<3, float, float> circ(std::make_pair(0.f, 2.f * 3.14159265358979f), util::get_range(cx), true);
        float err = 0;
        for (std::size_t i = 0; i < 1024; ++i) {
            const float th = 2.f * 3.14159265358979f * float(i) / 1024;
            err = std::max(err, std::abs(circ(th) - std::cos(th)));
        }
        expect(err < 2e-4f, "float periodic cubic on cos", err);
    }
    // vector-valued T (interpolation-test.cpp:674-703: a circle as Vec<2, float>): carried as two fields
    {
        struct Vec2f { float x, y; };
        constexpr std::size_t n = 31;
        const float two_pi = 2.f * 3.14159265358979f;
        std::vector<Vec2f> pts;
        std::vector<float> xs, ys;
        for (std::size_t i = 0; i < n; ++i) {
            pts.push_back({std::cos(two_pi * float(i) / n), std::sin(two_pi * float(i) / n)});
            xs.push_back(pts.back().x); ys.push_back(pts.back().y);
        }
        InterpolationFunction1D<3, Vec2f, float> circ(std::make_pair(0.f, two_pi), util::get_range(pts), true);
        InterpolationFunction1D<3, float, float> cx(std::make_pair(0.f, two_pi), util::get_range(xs), true);
        InterpolationFunction1D<3, float, float> cy(std::make_pair(0.f, two_pi), util::get_range(ys), true);
        float err = 0;
        bool same = true;
        std::vector<std::array<float, 1>> q;
        for (std::size_t i = 0; i < 257; ++i) {
            const float th = two_pi * float(i) / 257;
            q.push_back({th});
            const Vec2f v = circ(th);
            err = std::max(err, std::max(std::abs(v.x - std::cos(th)), std::abs(v.y - std::sin(th))));
            same = same && v.x == cx(th) && v.y == cy(th);
            const Vec2f d = circ.derivative({th}, 1);
            same = same && d.x == cx.derivative({th}, 1) && d.y == cy.derivative({th}, 1);
        }
        std::vector<Vec2f> batch;
        circ.evaluate(q, batch);
        for (std::size_t i = 0; i < q.size(); ++i) same = same && batch[i].x == cx(q[i][0]) && batch[i].y == cy(q[i][0]);
        const auto ctrl = circ.control_points();
        const auto ctrl_x = cx.control_points();
        for (std::size_t i = 0; i < n; ++i) same = same && ctrl(i).x == ctrl_x(i);
        expect(err < 2e-4f && same && decltype(circ)::components == 2, "Vec2f periodic cubic circle (two fields)", err);
    }
    std::printf("%d failure(s)\n", failures);
    return failures;
}
