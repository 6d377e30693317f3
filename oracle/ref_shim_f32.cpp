// TEST INFRASTRUCTURE ONLY.  The reference instantiated with T = U = float
// (InterpolationFunction<float, D, O, float>) for a few (D, O), to generate the
// fp32 golden fixture tests/golden/ref_outputs_f32.npz.  Same rules as
// ref_shim.cpp: compiled from the reference's headers where they lie, nothing
// copied.  Built on demand by tests/golden/make_ref_outputs_f32.py only.
#include <Interpolation.hpp>

#include <cstdint>
#include <vector>

#define SHIM_API extern "C" __attribute__((visibility("default")))

namespace {

struct Base {
    virtual ~Base() = default;
    virtual void eval(const float* pts, size_t q, float* out) const = 0;
    virtual void deriv(const float* pts, size_t q, const int* d, float* out) const = 0;
    virtual void spans(const float* pts, size_t q, int64_t* out) const = 0;
    virtual size_t ctrl(float* out) const = 0;
};

template <size_t D, size_t O>
struct RefF final : Base {
    intp::InterpolationFunction<float, D, O, float> fn;
    template <size_t... I>
    RefF(intp::util::index_sequence<I...>, const size_t* n, const int* per, const float* lo, const float* hi,
         const float* f) {
        std::array<size_t, D> dims;
        std::array<bool, D> p;
        for (size_t d = 0; d < D; ++d) { dims[d] = n[d]; p[d] = per[d] != 0; }
        intp::Mesh<float, D> mesh{intp::MeshDimension<D>(dims)};
        std::copy(f, f + mesh.size(), const_cast<float*>(mesh.data()));
        fn = intp::InterpolationFunction<float, D, O, float>(p, mesh, std::make_pair(lo[I], hi[I])...);
    }
    void eval(const float* pts, size_t q, float* out) const override {
        for (size_t i = 0; i < q; ++i) {
            std::array<float, D> c;
            for (size_t d = 0; d < D; ++d) c[d] = pts[i * D + d];
            out[i] = fn(c);
        }
    }
    void deriv(const float* pts, size_t q, const int* dv, float* out) const override {
        std::array<size_t, D> k;
        for (size_t d = 0; d < D; ++d) k[d] = size_t(dv[d]);
        for (size_t i = 0; i < q; ++i) {
            std::array<float, D> c;
            for (size_t d = 0; d < D; ++d) c[d] = pts[i * D + d];
            out[i] = fn.derivative(c, k);
        }
    }
    void spans(const float* pts, size_t q, int64_t* out) const override {
        const auto& sp = fn.spline();
        for (size_t i = 0; i < q; ++i)
            for (size_t d = 0; d < D; ++d) {
                float x = pts[i * D + d];
                out[i * D + d] = int64_t(sp.get_knot_iter(d, x, O) - sp.knots_begin(d)) - int64_t(O);
            }
    }
    size_t ctrl(float* out) const override {
        const auto& cp = fn.spline().control_points();
        if (out) std::copy(cp.begin(), cp.end(), out);
        return cp.size();
    }
};

template <size_t D, size_t O>
Base* make(const size_t* n, const int* per, const float* lo, const float* hi, const float* f) {
    return new RefF<D, O>(intp::util::make_index_sequence<D>{}, n, per, lo, hi, f);
}

}  // namespace

SHIM_API void* intp_ref32_create(int dim, int order, const uint64_t* n, const int* per, const float* lo,
                                 const float* hi, const float* f) {
    size_t nn[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) nn[d] = size_t(n[d]);
    try {
        if (dim == 1 && order == 3) return make<1, 3>(nn, per, lo, hi, f);
        if (dim == 1 && order == 5) return make<1, 5>(nn, per, lo, hi, f);
        if (dim == 2 && order == 3) return make<2, 3>(nn, per, lo, hi, f);
        if (dim == 2 && order == 2) return make<2, 2>(nn, per, lo, hi, f);
        if (dim == 3 && order == 3) return make<3, 3>(nn, per, lo, hi, f);
        if (dim == 3 && order == 1) return make<3, 1>(nn, per, lo, hi, f);
    } catch (const std::exception&) {
    }
    return nullptr;
}
SHIM_API void intp_ref32_destroy(void* h) { delete static_cast<Base*>(h); }
SHIM_API void intp_ref32_eval(void* h, const float* pts, uint64_t q, float* out) { static_cast<Base*>(h)->eval(pts, q, out); }
SHIM_API void intp_ref32_deriv(void* h, const float* pts, uint64_t q, const int* d, float* out) {
    static_cast<Base*>(h)->deriv(pts, q, d, out);
}
SHIM_API void intp_ref32_spans(void* h, const float* pts, uint64_t q, int64_t* out) { static_cast<Base*>(h)->spans(pts, q, out); }
SHIM_API uint64_t intp_ref32_ctrl(void* h, float* out) { return static_cast<Base*>(h)->ctrl(out); }
