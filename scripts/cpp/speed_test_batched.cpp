// The workload of the reference's interpolation-speed-test.cpp (:56-123: axes alternately periodic /
// non-periodic on [-pi, pi], field prod_d cos(i_d dt_d - pi), 2^20 uniform-random evaluation points,
// meshes of 2^p points) driven through the drop-in header the way a GPU wants it: one construction
// and ONE batched evaluate(points, out) per case instead of 2^20 single-point calls.  Host buffers,
// wall-clock (std::chrono) around the calls as the reference's Timer does, i.e. end to end
// including every host<->device copy.  Prints the table of BASELINE.md section 2.1.
//
// Build and run: scripts/run_reftests.sh
#define INTP_PERIODIC_NO_DUMMY_POINT  // the reference's own test configuration (test/CMakeLists.txt:46)
#include <intp_b200/Interpolation.hpp>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using namespace intp;
using clk = std::chrono::steady_clock;
static double ms(clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); }

constexpr std::size_t kEval = std::size_t{1} << 20;
constexpr double kPi = 3.14159265358979323846;

struct Row { double construct_ms, eval_ms, max_err; };

template <std::size_t D, std::size_t O, std::size_t... I>
Row run(std::size_t p, std::index_sequence<I...>) {
    const std::array<std::size_t, D> n{(std::size_t{1} << ((p + I) / D))...};
    const std::array<double, D> dt{(2 * kPi / double(n[I]))...};
    Mesh<double, D> mesh{MeshDimension<D>(n)};
    for (std::size_t i = 0; i < mesh.size(); ++i) {
        const auto idx = mesh.dimension().dimwise_indices(i);
        mesh(idx) = (... * std::cos(double(idx[I]) * dt[I] - kPi));
    }
    std::mt19937_64 gen(12345);
    std::uniform_real_distribution<> uni(-kPi, kPi);
    std::vector<std::array<double, D>> pts(kEval);
    for (auto& x : pts)
        for (auto& c : x) c = uni(gen);
    std::vector<double> out;

    const auto t0 = clk::now();
    InterpolationFunction<double, D, O> f({(I % 2 == 0)...}, mesh, ((void)I, std::make_pair(-kPi, kPi))...);
    const auto t1 = clk::now();
    f.evaluate(pts, out);
    const auto t2 = clk::now();

    // the sampled field: cos(x) on periodic axes; on non-periodic ones the n samples of cos(i dt - pi)
    // are spread over [-pi, pi] with spacing 2 pi / (n - 1)
    double err = 0;
    for (std::size_t i = 0; i < kEval; i += 997) {
        double ref = 1;
        for (std::size_t d = 0; d < D; ++d)
            ref *= d % 2 == 0 ? std::cos(pts[i][d]) : std::cos((pts[i][d] + kPi) * double(n[d] - 1) / double(n[d]) - kPi);
        err = std::max(err, std::abs(out[i] - ref));
    }
    return {ms(t0, t1), ms(t1, t2), err};
}

template <std::size_t D, std::size_t O>
void line(std::size_t p) {
    run<D, O>(p, std::make_index_sequence<D>{});  // first call pays one-off allocations
    Row r[3];
    for (Row& x : r) x = run<D, O>(p, std::make_index_sequence<D>{});
    // median of three, per column
    auto med = [&](double Row::*m) {
        double v[3] = {r[0].*m, r[1].*m, r[2].*m};
        std::sort(v, v + 3);
        return v[1];
    };
    const Row out{med(&Row::construct_ms), med(&Row::eval_ms), r[0].max_err};
    std::printf("| 2^%-2zu | %zu-D o%zu | %10.2f | %10.2f | %8.1f | %.1e |\n", p, D, O, out.construct_ms, out.eval_ms,
                double(kEval) / out.eval_ms * 1e-3, out.max_err);
    std::fflush(stdout);
}

int main() {
    std::printf("reference speed-test workload, batched through intp_b200 (host buffers, wall clock)\n");
    std::printf("| mesh | case | construct ms | evaluate 2^20 pts ms | Mpts/s | max |f - cos..| |\n|---|---|---|---|---|---|\n");
    // a first tiny case absorbs CUDA context creation
    run<1, 3>(12, std::make_index_sequence<1>{});
    const bool latency_only = std::getenv("SPEED_TEST_LATENCY_ONLY") != nullptr;
    for (std::size_t p : {20, 22, 24}) {
        if (latency_only) break;
        line<1, 3>(p); line<1, 5>(p);
        line<2, 3>(p); line<2, 5>(p);
        line<3, 3>(p); line<3, 5>(p);
    }
    {   // what a single-point operator() costs through the device (one launch + two tiny copies)
        Mesh<double, 3> m{64, 64, 64};
        for (std::size_t i = 0; i < m.size(); ++i) m.data()[i] = std::sin(1e-3 * double(i));
        InterpolationFunction<double, 3, 3> f(m, std::make_pair(0., 1.), std::make_pair(0., 1.), std::make_pair(0., 1.));
        double sink = f(.5, .5, .5);
        const int reps = 2000;
        const auto t0 = clk::now();
        for (int i = 0; i < reps; ++i) sink += f(.1 + 4e-4 * i, .5, .25);
        const auto t1 = clk::now();
        auto proxy = f.eval_proxy({.3, .4, .5});
        for (int i = 0; i < reps; ++i) sink += proxy(f);
        const auto t2 = clk::now();
        std::printf("\nsingle-point operator(): %.1f us per call; eval_proxy call: %.1f us (checksum %.6f)\n",
                    1e3 * ms(t0, t1) / reps, 1e3 * ms(t1, t2) / reps, sink);
    }
    return 0;
}
