// The reference's DEFAULT periodic convention (INTP_PERIODIC_NO_DUMMY_POINT undefined, README.md:57,
// InterpolationTemplate.hpp:255-265, :464-487): the last sample of a periodic axis is the dummy copy
// of the first one and is dropped.  Through the header that must equal what the C ABI gives for the
// same data without the dummy sample (n = N - 1 data points, same range).  Compiled WITHOUT the macro.
#include <intp_b200/Interpolation.hpp>

#include <cmath>
#include <cstdio>
#include <vector>

#ifdef INTP_PERIODIC_NO_DUMMY_POINT
#error "this test covers the default (dummy point) convention"
#endif

using namespace intp;

static int failures = 0;
static void expect(bool ok, const char* what) {
    std::printf("%-64s %s\n", what, ok ? "ok" : "FAILED");
    if (!ok) ++failures;
}

// the same spline straight from the C ABI: n data points per axis, no dummy sample anywhere
template <std::size_t D>
static std::vector<double> abi_values(int order, const int64_t* n, const int* per, const double* lo, const double* hi,
                                      const double* const* coords, const std::vector<double>& data,
                                      const std::vector<std::array<double, D>>& pts) {
    bspl_template* t = nullptr;
    bspl_function* f = nullptr;
    std::vector<double> out(pts.size());
    if (bspl_template_create(BSPL_F64, int(D), order, n, per, lo, hi, coords, 0, &t) != BSPL_OK ||
        bspl_template_interpolate(t, data.data(), 1, 0, nullptr, &f) != BSPL_OK ||
        bspl_evaluate(f, 0, pts.data(), int64_t(pts.size()), nullptr, out.data(), 0, nullptr) != BSPL_OK) {
        std::printf("C ABI: %s\n", bspl_last_error());
        ++failures;
    }
    bspl_function_destroy(f);
    bspl_template_destroy(t);
    return out;
}

int main() {
    // 1-D periodic quartic: 13 samples of which the last repeats the first; period = range
    {
        constexpr std::size_t N = 13;
        std::vector<double> f(N);
        for (std::size_t i = 0; i + 1 < N; ++i) f[i] = std::sin(0.5 * double(i)) + 0.1 * double(i % 3);
        f[N - 1] = f[0];
        InterpolationFunction<double, 1, 4> g(true, util::get_range(f), std::make_pair(0., 6.));
        std::vector<std::array<double, 1>> pts;
        for (int i = 0; i < 50; ++i) pts.push_back({-7. + 0.37 * i});
        const int64_t n[1] = {N - 1};
        const int per[1] = {1};
        const double lo[1] = {0.}, hi[1] = {6.};
        const auto ref = abi_values<1>(4, n, per, lo, hi, nullptr, std::vector<double>(f.begin(), f.end() - 1), pts);
        bool same = true;
        for (std::size_t i = 0; i < pts.size(); ++i) same = same && g(pts[i]) == ref[i];
        expect(same, "1-D periodic: dummy sample dropped, period = range");
        expect(std::abs(g(6.) - g(0.)) < 1e-14 && std::abs(g(0.5 * 3) - f[3]) < 1e-13, "1-D periodic: closes and interpolates the samples");
        expect(g.control_points().size() == N - 1, "1-D periodic: N - 1 control points");
        // default x range of the 1-D convenience class: [0, N - 1] whether periodic or not
        InterpolationFunction1D<3> d1(util::get_range(f), true);
        expect(d1.range(0).first == 0. && d1.range(0).second == double(N - 1), "InterpolationFunction1D default range [0, N-1]");
    }
    // 2-D, x periodic (dummy row), y not; through a template, two fields
    {
        constexpr std::size_t NX = 9, NY = 7;
        Mesh<double, 2> a(NX, NY), b(NX, NY);
        for (std::size_t i = 0; i < NX; ++i)
            for (std::size_t j = 0; j < NY; ++j) {
                const std::size_t ii = i % (NX - 1);  // row NX-1 repeats row 0
                a(i, j) = std::cos(0.8 * double(ii)) * (1. + 0.2 * double(j));
                b(i, j) = double(ii * j) - 3.;
            }
        InterpolationFunctionTemplate<double, 2, 3> tm({true, false}, a.dimension(), std::make_pair(-1., 1.), std::make_pair(0., 3.));
        auto ga = tm.interpolate(a);
        InterpolationFunction<double, 2, 3> gb;
        tm.interpolate(gb, b);
        std::vector<std::array<double, 2>> pts;
        for (int i = 0; i < 60; ++i) pts.push_back({-1.5 + 0.05 * i, 0.05 * i});
        std::vector<double> kept;
        for (std::size_t i = 0; i + 1 < NX; ++i)
            for (std::size_t j = 0; j < NY; ++j) kept.push_back(b(i, j));
        const int64_t n[2] = {NX - 1, NY};
        const int per[2] = {1, 0};
        const double lo[2] = {-1., 0.}, hi[2] = {1., 3.};
        const auto ref = abi_values<2>(3, n, per, lo, hi, nullptr, kept, pts);
        bool same = true;
        for (std::size_t i = 0; i < pts.size(); ++i) same = same && gb(pts[i]) == ref[i];
        expect(same, "2-D (periodic, non-periodic) through a template");
        expect(std::abs(ga(-1. + 0.25 * 3, 1.) - a(3, 2)) < 1e-13, "2-D: interpolates the samples");
        const auto ctrl = gb.control_points();
        expect(ctrl.dim_size(0) == NX - 1 && ctrl.dim_size(1) == NY, "2-D: control points (NX - 1) x NY");
    }
    // non-uniform periodic axis: N abscissae for N samples (the last one closes the period)
    {
        const std::vector<double> x{0., 0.4, 1.1, 1.5, 2.3, 3.0, 3.2, 4.0};
        std::vector<double> f(x.size());
        for (std::size_t i = 0; i + 1 < x.size(); ++i) f[i] = std::sin(2. * 3.14159265358979 * x[i] / 4.0) + 0.3 * double(i % 2);
        f.back() = f.front();
        InterpolationFunction<double, 1, 3> g(true, util::get_range(f), util::get_range(x));
        std::vector<std::array<double, 1>> pts;
        for (int i = 0; i < 40; ++i) pts.push_back({-1. + 0.17 * i});
        const int64_t n[1] = {int64_t(x.size()) - 1};
        const int per[1] = {1};
        const double lo[1] = {0.}, hi[1] = {4.};
        const double* coords[1] = {x.data()};
        const auto ref = abi_values<1>(3, n, per, lo, hi, coords, std::vector<double>(f.begin(), f.end() - 1), pts);
        bool same = true;
        for (std::size_t i = 0; i < pts.size(); ++i) same = same && g(pts[i]) == ref[i];
        expect(same && !g.uniform(0), "1-D non-uniform periodic");
        expect(std::abs(g(x[4]) - f[4]) < 1e-13, "1-D non-uniform periodic: interpolates the samples");
    }
    std::printf("%d failure(s)\n", failures);
    return failures;
}
