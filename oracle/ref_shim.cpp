// TEST INFRASTRUCTURE ONLY -- never linked, imported or called by the product path.
//
// C shim over the UNMODIFIED reference headers (found at build time with
// -I$(REF)/src/include, never copied into this repo).  It pre-instantiates
// intp::InterpolationFunction<double, D, O> for D in 1..3, O in 0..5 and
// exposes what the parity tests and the CPU-baseline timing need: knots,
// ranges, spans, plain control points, values, mixed derivatives and a timed
// InterpolationFunctionTemplate::interpolate.  Three shared objects are built
// from this one file (see oracle/Makefile):
//   _ref/libintp_ref_cell.so   -DINTP_CELL_LAYOUT -DINTP_MULTITHREAD  (the
//                              reference's own Release test configuration,
//                              test/CMakeLists.txt:43-47) -> eval + timing
//   _ref/libintp_ref_plain.so  no INTP_CELL_LAYOUT -> spline().control_points()
//                              is the plain N-d array (BSpline.hpp:58-61)
//   _ref/libintp_ref_plain_mt.so  plain layout + INTP_MULTITHREAD -> the threaded
//                              solve baseline without the 16x cell-layout fill
// All define INTP_PERIODIC_NO_DUMMY_POINT like the reference test build.
#include <Interpolation.hpp>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define SHIM_API extern "C" __attribute__((visibility("default")))

namespace {

struct AxisSpec {
    int periodic;
    double lo, hi;          // uniform range
    const double* coords;   // non-NULL -> non-uniform abscissae, n_coords entries
    size_t n_coords;
};

struct RefBase {
    virtual ~RefBase() = default;
    virtual void eval(const double* pts, size_t q, double* out) const = 0;
    virtual void deriv(const double* pts, size_t q, const int* d, double* out) const = 0;
    virtual void spans(const double* pts, size_t q, int64_t* out) const = 0;
    virtual size_t knots(int axis, double* out) const = 0;
    virtual void range(int axis, double* lo_hi) const = 0;
    virtual size_t ctrl_size() const = 0;
    virtual void ctrl(double* out) const = 0;
    virtual int dim() const = 0;
    double t_template_ms = 0, t_interpolate_ms = 0;
};

using clk = std::chrono::steady_clock;
inline double ms_since(clk::time_point t0) {
    return std::chrono::duration<double, std::milli>(clk::now() - t0).count();
}

template <size_t D, size_t O>
struct Ref final : RefBase {
    using Fn = intp::InterpolationFunction<double, D, O, double>;
    using Tm = intp::InterpolationFunctionTemplate<double, D, O, double>;
    Fn fn;

    // Mask bit d set -> axis d non-uniform (iterator-pair overload,
    // Interpolation.hpp:365-464); clear -> uniform (value-pair overload, :322-362).
    template <size_t Mask, size_t... I>
    static Tm make_template(intp::util::index_sequence<I...>,
                            const std::array<bool, D>& per,
                            const intp::MeshDimension<D>& md,
                            const AxisSpec* ax) {
        return Tm(per, md, axis_arg<(Mask >> I) & 1>(ax[I])...);
    }
    template <size_t NonUniform>
    static typename std::conditional<NonUniform != 0,
                                     std::pair<const double*, const double*>,
                                     std::pair<double, double>>::type
    axis_arg(const AxisSpec& a) {
        if constexpr (NonUniform != 0) {
            return std::make_pair(a.coords, a.coords + a.n_coords);
        } else {
            return std::make_pair(a.lo, a.hi);
        }
    }

    template <size_t Mask>
    static bool try_build(size_t mask, Ref& self, const std::array<bool, D>& per,
                          const intp::Mesh<double, D>& mesh, const AxisSpec* ax,
                          int repeat) {
        if (mask != Mask) return false;
        auto t0 = clk::now();
        Tm tm = make_template<Mask>(intp::util::make_index_sequence<D>{}, per,
                                    mesh.dimension(), ax);
        self.t_template_ms = ms_since(t0);
        double best = 1e300;
        for (int r = 0; r < (repeat < 1 ? 1 : repeat); ++r) {
            t0 = clk::now();
            self.fn = tm.interpolate(mesh);
            best = std::min(best, ms_since(t0));
        }
        self.t_interpolate_ms = best;
        return true;
    }

    Ref(const size_t* n, const AxisSpec* ax, const double* f, int repeat) {
        std::array<size_t, D> dims;
        std::array<bool, D> per;
        size_t mask = 0;
        for (size_t d = 0; d < D; ++d) {
            dims[d] = n[d];
            per[d] = ax[d].periodic != 0;
            if (ax[d].coords) mask |= size_t{1} << d;
        }
        intp::Mesh<double, D> mesh{intp::MeshDimension<D>(dims)};
        // Mesh exposes only a const data(); the storage is row-major like f (Mesh.hpp:241-246)
        std::copy(f, f + mesh.size(), const_cast<double*>(mesh.data()));
        bool ok = try_build<0>(mask, *this, per, mesh, ax, repeat);
        if constexpr (D == 1) {
            ok = ok || try_build<1>(mask, *this, per, mesh, ax, repeat);
        } else if constexpr (D == 2) {
            ok = ok || try_build<1>(mask, *this, per, mesh, ax, repeat) ||
                 try_build<2>(mask, *this, per, mesh, ax, repeat) ||
                 try_build<3>(mask, *this, per, mesh, ax, repeat);
        } else if constexpr (D == 3) {
            ok = ok || try_build<7>(mask, *this, per, mesh, ax, repeat) ||
                 try_build<4>(mask, *this, per, mesh, ax, repeat);
        }
        if (!ok) throw std::runtime_error("unsupported uniform/non-uniform mix");
    }

    int dim() const override { return int(D); }

    void eval(const double* pts, size_t q, double* out) const override {
        for (size_t i = 0; i < q; ++i) {
            std::array<double, D> c;
            for (size_t d = 0; d < D; ++d) c[d] = pts[i * D + d];
            out[i] = fn(c);
        }
    }
    void deriv(const double* pts, size_t q, const int* dv, double* out) const override {
        std::array<size_t, D> k;
        for (size_t d = 0; d < D; ++d) k[d] = size_t(dv[d]);
        for (size_t i = 0; i < q; ++i) {
            std::array<double, D> c;
            for (size_t d = 0; d < D; ++d) c[d] = pts[i * D + d];
            out[i] = fn.derivative(c, k);
        }
    }
    // span - order per axis through the reference's own get_knot_iter with a
    // deliberately useless hint (= order), i.e. the upper_bound branch or the
    // hint-accept branch of BSpline.hpp:146-156, whichever the reference takes.
    void spans(const double* pts, size_t q, int64_t* out) const override {
        const auto& sp = fn.spline();
        for (size_t i = 0; i < q; ++i) {
            for (size_t d = 0; d < D; ++d) {
                double x = pts[i * D + d];
                auto it = sp.get_knot_iter(d, x, O);
                out[i * D + d] = int64_t(it - sp.knots_begin(d)) - int64_t(O);
            }
        }
    }
    size_t knots(int axis, double* out) const override {
        const auto& sp = fn.spline();
        size_t k = sp.knots_num(size_t(axis));
        if (out) std::copy(sp.knots_begin(size_t(axis)), sp.knots_end(size_t(axis)), out);
        return k;
    }
    void range(int axis, double* lo_hi) const override {
        lo_hi[0] = fn.range(size_t(axis)).first;
        lo_hi[1] = fn.range(size_t(axis)).second;
    }
    size_t ctrl_size() const override { return fn.spline().control_points().size(); }
    void ctrl(double* out) const override {
        const auto& cp = fn.spline().control_points();
        std::copy(cp.begin(), cp.end(), out);
    }
};

template <size_t D>
RefBase* make_order(int order, const size_t* n, const AxisSpec* ax, const double* f, int rep) {
    switch (order) {
        case 0: return new Ref<D, 0>(n, ax, f, rep);
        case 1: return new Ref<D, 1>(n, ax, f, rep);
        case 2: return new Ref<D, 2>(n, ax, f, rep);
        case 3: return new Ref<D, 3>(n, ax, f, rep);
        case 4: return new Ref<D, 4>(n, ax, f, rep);
        case 5: return new Ref<D, 5>(n, ax, f, rep);
        case 6: return new Ref<D, 6>(n, ax, f, rep);
        case 7: return new Ref<D, 7>(n, ax, f, rep);
        default: return nullptr;
    }
}

template <typename F>
void split_threads(size_t q, int nthreads, F&& body) {
    if (nthreads <= 1 || q < 1024) { body(size_t{0}, q); return; }
    std::vector<std::thread> th;
    size_t chunk = (q + size_t(nthreads) - 1) / size_t(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        size_t b = std::min(q, size_t(t) * chunk), e = std::min(q, b + chunk);
        if (b < e) th.emplace_back([=, &body] { body(b, e); });
    }
    for (auto& t : th) t.join();
}

}  // namespace

// periodic[D], lo[D], hi[D]; coords[d] may be NULL (uniform axis) or point to
// n_coords[d] abscissae.  f is the row-major mesh.  repeat >= 1 re-runs
// interpolate() and keeps the best wall time (for the CPU baseline).
SHIM_API void* intp_ref_create(int dim, int order, const uint64_t* n, const int* periodic,
                               const double* lo, const double* hi,
                               const double* const* coords, const uint64_t* n_coords,
                               const double* f, int repeat) {
    try {
        AxisSpec ax[4];
        size_t nn[4];
        if (dim < 1 || dim > 4) return nullptr;
        for (int d = 0; d < dim; ++d) {
            nn[d] = size_t(n[d]);
            ax[d] = AxisSpec{periodic[d], lo[d], hi[d],
                             coords ? coords[d] : nullptr,
                             (coords && coords[d]) ? size_t(n_coords[d]) : 0};
        }
        switch (dim) {
            case 1: return make_order<1>(order, nn, ax, f, repeat);
            case 2: return make_order<2>(order, nn, ax, f, repeat);
            case 3: return make_order<3>(order, nn, ax, f, repeat);
            case 4: return make_order<4>(order, nn, ax, f, repeat);
            default: return nullptr;
        }
    } catch (const std::exception&) {
        return nullptr;
    }
}
SHIM_API void intp_ref_destroy(void* h) { delete static_cast<RefBase*>(h); }

SHIM_API void intp_ref_eval(void* h, const double* pts, uint64_t q, double* out, int nthreads) {
    auto* r = static_cast<RefBase*>(h);
    size_t D = size_t(r->dim());
    split_threads(size_t(q), nthreads,
                  [&](size_t b, size_t e) { r->eval(pts + b * D, e - b, out + b); });
}
SHIM_API void intp_ref_deriv(void* h, const double* pts, uint64_t q, const int* d, double* out,
                             int nthreads) {
    auto* r = static_cast<RefBase*>(h);
    size_t D = size_t(r->dim());
    split_threads(size_t(q), nthreads,
                  [&](size_t b, size_t e) { r->deriv(pts + b * D, e - b, d, out + b); });
}
SHIM_API void intp_ref_spans(void* h, const double* pts, uint64_t q, int64_t* out) {
    static_cast<RefBase*>(h)->spans(pts, size_t(q), out);
}
SHIM_API uint64_t intp_ref_knots(void* h, int axis, double* out) {
    return static_cast<RefBase*>(h)->knots(axis, out);
}
SHIM_API void intp_ref_range(void* h, int axis, double* lo_hi) {
    static_cast<RefBase*>(h)->range(axis, lo_hi);
}
SHIM_API uint64_t intp_ref_ctrl_size(void* h) { return static_cast<RefBase*>(h)->ctrl_size(); }
SHIM_API void intp_ref_ctrl(void* h, double* out) { static_cast<RefBase*>(h)->ctrl(out); }
SHIM_API double intp_ref_time_template_ms(void* h) { return static_cast<RefBase*>(h)->t_template_ms; }
SHIM_API double intp_ref_time_interpolate_ms(void* h) {
    return static_cast<RefBase*>(h)->t_interpolate_ms;
}
SHIM_API int intp_ref_cell_layout(void) {
#ifdef INTP_CELL_LAYOUT
    return 1;
#else
    return 0;
#endif
}
SHIM_API int intp_ref_multithread(void) {
#ifdef INTP_MULTITHREAD
    return 1;
#else
    return 0;
#endif
}

// BandLU on a dense n x n row-major matrix `a` (entries outside the band /
// the cyclic corners are ignored): factor with the reference and solve one
// right-hand side in place (band-matrix-and-solver-test.cpp:11-32).
SHIM_API int intp_ref_band_solve(uint64_t n, uint64_t p, uint64_t q, int cyclic, const double* a,
                                 double* x) {
    using namespace intp;
    try {
        if (!cyclic) {
            BandMatrix<double> m{size_t(n), size_t(p), size_t(q)};
            for (size_t i = 0; i < n; ++i)
                for (size_t j = (i > p ? i - p : 0); j < std::min<size_t>(n, i + q + 1); ++j)
                    m(i, j) = a[i * n + j];
            BandLU<BandMatrix<double>> lu{std::move(m)};
            lu.solve_in_place(x);
        } else {
            ExtendedBandMatrix<double> m{size_t(n), size_t(p), size_t(q)};
            for (size_t i = 0; i < n; ++i) {
                for (size_t j = 0; j < n; ++j) {
                    bool in_main = (j + p >= i) && (i + q >= j);
                    bool in_right = j >= std::max<size_t>(n - p, i + q + 1);
                    bool in_bottom = i >= std::max<size_t>(n - q, j + p + 1);
                    if (in_main || in_right || in_bottom) m(i, j) = a[i * n + j];
                }
            }
            BandLU<ExtendedBandMatrix<double>> lu{std::move(m)};
            lu.solve_in_place(x);
        }
        return 0;
    } catch (const std::exception&) {
        return 1;
    }
}
