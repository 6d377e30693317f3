// BandMatrix / ExtendedBandMatrix / BandLU of include/intp_b200 in the scenarios of the reference's
// band-matrix-and-solver-test.cpp (:52-115): Laplacian, the circulant collocation matrices of the
// quadratic and quartic periodic splines, and an uneven cyclic band; residual ||A x - b|| / ||b||
// below the reference's 1e-10 (:30).  Links either libbspline_b200.so (GPU) or band_rows_stub.cpp
// (CPU suite).
#include <BandLU.hpp>
#include <BandMatrix.hpp>

#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

using namespace intp;

static int failures = 0;

template <typename Mat, typename Rhs>
void check_solver(const char* name, Mat mat, Rhs&& b) {
    const std::size_t n = mat.dim();
    std::vector<double> b_copy(n);
    for (std::size_t i = 0; i < n; ++i) b_copy[i] = b[i];
    BandLU<Mat> solver{mat};
    auto x = solver.solve(b);
    std::vector<double> xv(n);
    for (std::size_t i = 0; i < n; ++i) xv[i] = x[i];
    const std::vector<double> ax = mat * xv;
    double err = 0, l2 = 0;
    for (std::size_t i = 0; i < n; ++i) { err += (ax[i] - b_copy[i]) * (ax[i] - b_copy[i]); l2 += b_copy[i] * b_copy[i]; }
    const double d = std::sqrt(err / l2);
    std::printf("%-44s ||b - A x|| / ||b|| = %.3e\n", name, d);
    if (!(d < 1e-10)) ++failures;
}

int main() {
    constexpr std::size_t n = 64;
    std::mt19937 gen(20240607u);
    std::uniform_real_distribution<> uni(-1., 1.);
    std::vector<double> b(n);
    for (auto& v : b) v = uni(gen);

    {
        BandMatrix<double> lap{n, 1, 1};
        for (std::size_t i = 0; i < n; ++i) {
            lap(i, i) = -2;
            if (i > 0) lap(i, i - 1) = 1;
            if (i + 1 < n) lap(i, i + 1) = 1;
        }
        std::vector<double> rhs(b);
        check_solver("band, Laplacian (pointer rhs, in place)", lap, rhs.data());
        bool threw = false;
        try { lap(0, 5) = 1; } catch (const std::out_of_range&) { threw = true; }
        if (!threw) { std::puts("write outside the band was accepted"); ++failures; }
    }
    {
        ExtendedBandMatrix<double> m{n, 1, 1};
        for (std::size_t i = 0; i < n; ++i) {
            m(i, i) = 3. / 4.;
            m(i, (i + n - 1) % n) = 1. / 8.;
            m(i, (i + 1) % n) = 1. / 8.;
        }
        check_solver("cyclic, quadratic B-spline circulant", m, b);
    }
    {
        ExtendedBandMatrix<double> m{n, 2, 2};
        for (std::size_t i = 0; i < n; ++i) {
            m(i, i) = 115. / 192.;
            m(i, (i + n - 1) % n) = m(i, (i + 1) % n) = 19. / 96.;
            m(i, (i + n - 2) % n) = m(i, (i + 2) % n) = 1. / 384.;
        }
        check_solver("cyclic, quartic B-spline circulant", m, b);
        const ExtendedBandMatrix<double>& cm = m;
        if (cm(0, n - 2) != 1. / 384. || cm(n - 1, 1) != 1. / 384. || cm(3, 4) != 19. / 96.) {
            std::puts("corner / band read-back wrong");
            ++failures;
        }
    }
    {
        ExtendedBandMatrix<double> m{n, 1, 3};
        for (std::size_t i = 0; i < n; ++i) {
            m(i, (i + n - 1) % n) = 2889. / 16000.;
            m(i, i) = 1701. / 3200.;
            m(i, (i + 1) % n) = 33. / 128.;
            m(i, (i + 2) % n) = 729. / 20000.;
            m(i, (i + 3) % n) = 9. / 20000.;
        }
        check_solver("cyclic, uneven band (p = 1, q = 3)", m, b);
    }
    std::printf("%s\n", failures ? "FAILED" : "all band solver checks passed");
    return failures ? 1 : 0;
}
