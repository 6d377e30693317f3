// intp_b200/Interpolation.hpp -- drop-in host header for the B200 implementation of the
// BSplineInterpolation hot path.  Same namespace (`intp`), class names and member
// signatures as the reference's user API:
//   Mesh / MeshDimension              <- src/include/Mesh.hpp:11-116, :125-383
//   InterpolationFunction{,1D}        <- src/include/Interpolation.hpp:17-540
//   InterpolationFunctionTemplate{,1D}<- src/include/InterpolationTemplate.hpp:32-604
// but every number is produced on the GPU through the C ABI of bspline_b200.h (link
// libbspline_b200.so).  INTP_PERIODIC_NO_DUMMY_POINT keeps its meaning (reference README.md:57):
// defined, a periodic axis with N samples has period N*dx and its closing sample is implicit;
// undefined (the reference's default), the last sample of a periodic axis is the dummy copy of the
// first one and is dropped (InterpolationTemplate.hpp:255-265, :464-487).  The other reference
// macros (INTP_CELL_LAYOUT, INTP_MULTITHREAD, allocators) describe CPU layout / threading and have
// no effect here.  Added: batched evaluate()/value_grad() on host or device pointers and a batched
// interpolate() for many fields.
//
// Requires C++17.  U (coordinates, arithmetic) is double or float; T is U, another arithmetic type
// (converted), or a trivially copyable aggregate of U values (carried as fields).
#ifndef INTP_B200_INTERPOLATION_HPP
#define INTP_B200_INTERPOLATION_HPP

#include <array>
#include <cstring>
#include <cstddef>
#include <cstdint>
#include <iterator>
#include <memory>
#include <new>
#include <numeric>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../bspline_b200.h"
#include "BSpline.hpp"
#include "Mesh.hpp"
#include "util.hpp"

namespace intp {

// ---------------------------------------------------------------- ABI plumbing
namespace b200_detail {

inline void check(int rc) {
    if (rc == BSPL_OK) return;
    const std::string msg = bspl_last_error();
    switch (rc) {
        case BSPL_ERR_DOMAIN: throw std::domain_error(msg);   // Interpolation.hpp:486
        case BSPL_ERR_ALLOC: throw std::bad_alloc();
        default: throw std::runtime_error(msg);               // INTP_ASSERT, util.hpp:247-258
    }
}

template <typename T> constexpr bspl_dtype dtype_of() {
    static_assert(std::is_same_v<T, double> || std::is_same_v<T, float>, "T must be double or float");
    return std::is_same_v<T, double> ? BSPL_F64 : BSPL_F32;
}

// How a value type T travels on a library whose arithmetic runs in the coordinate type U:
//  - T == U: directly;
//  - another arithmetic T (the reference's InterpolationFunction<float, D, O> on double coordinates):
//    one field, converted on the way in and out -- the arithmetic is U's, only the storage was T's;
//  - vector-valued T (e.g. the reference's Vec<2, float> circle, interpolation-test.cpp:674-703): any
//    trivially copyable aggregate of K values of U is carried as K fields of one function -- K
//    independent scalar splines sharing knots, factors and query work.
// the scalar an aggregate says it is made of (T::value_type), or U when it does not say
template <typename T, typename U, typename = void>
struct aggregate_scalar { using type = U; };
template <typename T, typename U>
struct aggregate_scalar<T, U, std::void_t<typename T::value_type>> { using type = typename T::value_type; };
template <typename T, typename U>
struct components_of {
    static_assert(std::is_arithmetic_v<T> || (std::is_trivially_copyable_v<T> && sizeof(T) % sizeof(U) == 0),
                  "T must be arithmetic or a trivially copyable aggregate of values of the coordinate type U");
    // Vec<2, float> on the default double coordinates would be reinterpreted as one double: the reference
    // computes such a T component-wise in T's own arithmetic, here the components must BE coordinates
    static_assert(std::is_arithmetic_v<T> || std::is_same_v<typename aggregate_scalar<T, U>::type, U>,
                  "a vector-valued T must consist of values of the coordinate type U "
                  "(e.g. InterpolationFunction<Vec<2, float>, D, O, float>)");
    static constexpr std::size_t value = std::is_arithmetic_v<T> ? 1 : sizeof(T) / sizeof(U);
};
// T is handed to the library as it is
template <typename T, typename U>
constexpr bool direct_v = std::is_same_v<T, U>;
// [m][K] interleaved -> [K][m]
template <typename T, typename U>
std::vector<U> split_components(const T* data, std::size_t m) {
    constexpr std::size_t K = components_of<T, U>::value;
    std::vector<U> out(K * m);
    if constexpr (std::is_arithmetic_v<T>) {
        for (std::size_t i = 0; i < m; ++i) out[i] = static_cast<U>(data[i]);
    } else {
        const U* in = reinterpret_cast<const U*>(data);
        for (std::size_t i = 0; i < m; ++i)
            for (std::size_t k = 0; k < K; ++k) out[k * m + i] = in[i * K + k];
    }
    return out;
}
// [m] values of component k -> slot k of [m][K] interleaved storage
template <typename T, typename U>
void merge_component(const U* field, std::size_t m, std::size_t k, T* data) {
    constexpr std::size_t K = components_of<T, U>::value;
    if constexpr (std::is_arithmetic_v<T>) {
        for (std::size_t i = 0; i < m; ++i) data[i] = static_cast<T>(field[i]);
    } else {
        U* out = reinterpret_cast<U*>(data);
        for (std::size_t i = 0; i < m; ++i) out[i * K + k] = field[i];
    }
}
// one value from its K components
template <typename T, typename U>
T from_components(const U* c) {
    if constexpr (std::is_arithmetic_v<T>) {
        return static_cast<T>(c[0]);
    } else {
        T v;
        std::memcpy(&v, c, sizeof(T));
        return v;
    }
}

struct FnDeleter { void operator()(bspl_function* p) const { bspl_function_destroy(p); } };
struct PlanDeleter { void operator()(bspl_query_plan* p) const { bspl_query_plan_destroy(p); } };
struct TmDeleter { void operator()(bspl_template* p) const { bspl_template_destroy(p); } };
// shared only with the BSpline view returned by spline(); copies of a function clone the device storage
using FnHandle = std::shared_ptr<bspl_function>;
inline FnHandle own(bspl_function* p) { return FnHandle(p, FnDeleter()); }
using TmHandle = std::unique_ptr<bspl_template, TmDeleter>;

// One axis of a template: a (min, max) pair of numbers -> uniform; a pair of iterators
// over the abscissae -> non-uniform (the two create_knot_vector_ overloads).
struct AxisSpec {
    double lo = 0, hi = 0;
    std::vector<double> coords;
};
template <typename X>
AxisSpec make_axis(const std::pair<X, X>& r) {
    AxisSpec a;
    if constexpr (std::is_arithmetic_v<X>) {
        a.lo = static_cast<double>(r.first);
        a.hi = static_cast<double>(r.second);
    } else {
        for (auto it = r.first; it != r.second; ++it) a.coords.push_back(static_cast<double>(*it));
        if (a.coords.size() < 2) throw std::runtime_error("an axis needs at least two coordinates");
        a.lo = a.coords.front();
        a.hi = a.coords.back();
    }
    return a;
}

#ifdef INTP_PERIODIC_NO_DUMMY_POINT
constexpr bool kDummyPoint = false;
#else
constexpr bool kDummyPoint = true;
#endif

// Drop the closing (dummy) sample of every periodic axis: [N0]..[N_{D-1}] -> [N_d - periodic_d].
template <typename T, std::size_t D>
std::vector<T> strip_dummy(const T* data, const MeshDimension<D>& full, const std::array<bool, D>& periodic) {
    typename MeshDimension<D>::index_type ext{};
    for (std::size_t d = 0; d < D; ++d) ext[d] = full.dim_size(d) - (periodic[d] ? 1 : 0);
    const MeshDimension<D> kept(ext);
    std::vector<T> out(kept.size());
    for (std::size_t i = 0; i < out.size(); ++i) out[i] = data[full.indexing(kept.dimwise_indices(i))];
    return out;
}

inline int& default_device() {
    static int dev = 0;
    return dev;
}

}  // namespace b200_detail

// Device used by objects constructed afterwards (CUDA ordinal; default 0).
inline void set_device(int ordinal) { b200_detail::default_device() = ordinal; }

template <typename T, std::size_t D, std::size_t O, typename U>
class InterpolationFunctionTemplate;
template <typename T, std::size_t D, std::size_t O, typename U>
class InterpolationFunction;

// What eval_proxy returns (Interpolation.hpp:493-506, InterpolationTemplate.hpp:145-176): the
// query-dependent work done once, applicable to any function of the same template.  Copyable.
template <typename T, std::size_t D, std::size_t O, typename U>
class EvalProxy {
   public:
    using function_type = InterpolationFunction<T, D, O, U>;
    EvalProxy(const bspl_function* fn, const U* points, std::size_t q) : q_(q) {
        bspl_query_plan* p = nullptr;
        b200_detail::check(bspl_query_plan_create(fn, points, static_cast<int64_t>(q), 0, nullptr, &p));
        plan_.reset(p, b200_detail::PlanDeleter());
    }
    // from the template alone, before any field is interpolated (InterpolationTemplate.hpp:145-165)
    EvalProxy(const bspl_template* tm, const U* points, std::size_t q) : q_(q) {
        bspl_query_plan* p = nullptr;
        b200_detail::check(bspl_template_query_plan_create(tm, points, static_cast<int64_t>(q), 0, nullptr, &p));
        plan_.reset(p, b200_detail::PlanDeleter());
    }
    // single point, like the reference's closure: proxy(interp) -> value
    T operator()(const function_type& interp) const {
        constexpr std::size_t K = b200_detail::components_of<T, U>::value;
        U buf[K];
        for (std::size_t k = 0; k < K; ++k)
            b200_detail::check(bspl_query_plan_evaluate(plan_.get(), interp.handle(), static_cast<int64_t>(k), nullptr, 0,
                                                        &buf[k], 0, nullptr));
        return b200_detail::from_components<T, U>(buf);
    }
    // batched: out[q]
    void operator()(const function_type& interp, T* out) const {
        constexpr std::size_t K = b200_detail::components_of<T, U>::value;
        if constexpr (b200_detail::direct_v<T, U>) {
            b200_detail::check(bspl_query_plan_evaluate(plan_.get(), interp.handle(), 0, nullptr, 0, out, 0, nullptr));
        } else {
            std::vector<U> tmp(q_);
            for (std::size_t k = 0; k < K; ++k) {
                b200_detail::check(bspl_query_plan_evaluate(plan_.get(), interp.handle(), static_cast<int64_t>(k), nullptr, 0,
                                                            tmp.data(), 0, nullptr));
                b200_detail::merge_component<T, U>(tmp.data(), q_, k, out);
            }
        }
    }
    void value_grad(const function_type& interp, T* out) const {
        static_assert(b200_detail::direct_v<T, U>, "value_grad: T must be the coordinate type");
        b200_detail::check(bspl_query_plan_evaluate(plan_.get(), interp.handle(), 0, nullptr, 1, out, 0, nullptr));
    }
    std::size_t size() const { return q_; }

   private:
    std::shared_ptr<bspl_query_plan> plan_;
    std::size_t q_;
};

// ---------------------------------------------------------------- InterpolationFunction
template <typename T, std::size_t D, std::size_t O, typename U = double>
class InterpolationFunction {
    static_assert(D >= 1 && D <= BSPL_MAX_DIM && O <= BSPL_MAX_ORDER, "dim 1..4, order 0..7 (the reference takes any; these are the instantiated device kernels)");

   public:
    using val_type = T;
    using coord_type = U;
    using size_type = std::size_t;
    static constexpr size_type dim = D;
    static constexpr size_type order = O;
    // scalar T: 1; vector-valued T: number of U values it holds (one field each)
    static constexpr size_type components = b200_detail::components_of<T, U>::value;
    template <typename V> using DimArray = std::array<V, D>;
    friend class InterpolationFunctionTemplate<T, D, O, U>;

    InterpolationFunction() = default;

    // (periodicity, mesh, ranges...)  Interpolation.hpp:92-101
    template <typename... Ts, typename = std::enable_if_t<sizeof...(Ts) == D>>
    InterpolationFunction(DimArray<bool> periodicity, const Mesh<T, D>& f_mesh, std::pair<Ts, Ts>... x_ranges)
        : InterpolationFunction(
              InterpolationFunctionTemplate<T, D, O, U>(periodicity, f_mesh.dimension(), x_ranges...).interpolate(f_mesh)) {}
    // all axes non-periodic  Interpolation.hpp:104-107
    template <typename... Ts, typename = std::enable_if_t<sizeof...(Ts) == D>>
    InterpolationFunction(const Mesh<T, D>& f_mesh, std::pair<Ts, Ts>... x_ranges)
        : InterpolationFunction(DimArray<bool>{}, f_mesh, x_ranges...) {}
    // 1-D iterator forms  Interpolation.hpp:49-79
    template <typename It, typename C1, typename C2,
              typename = std::enable_if_t<D == 1 && std::is_convertible_v<typename std::iterator_traits<It>::iterator_category,
                                                                           std::input_iterator_tag>>>
    InterpolationFunction(bool periodic, std::pair<It, It> f_range, std::pair<C1, C2> x_range)
        : InterpolationFunction(DimArray<bool>{periodic}, Mesh<T, 1>(f_range),
                                std::pair<std::common_type_t<C1, C2>, std::common_type_t<C1, C2>>(x_range)) {}
    template <typename It, typename C1, typename C2,
              typename = std::enable_if_t<D == 1 && std::is_convertible_v<typename std::iterator_traits<It>::iterator_category,
                                                                           std::input_iterator_tag>>>
    InterpolationFunction(std::pair<It, It> f_range, std::pair<C1, C2> x_range)
        : InterpolationFunction(false, f_range, x_range) {}

    // value semantics, like the reference (device storage is cloned)
    InterpolationFunction(const InterpolationFunction& o) { *this = o; }
    InterpolationFunction& operator=(const InterpolationFunction& o) {
        if (this != &o) {
            h_.reset();
            if (o.h_) {
                bspl_function* p = nullptr;
                b200_detail::check(bspl_function_clone(o.h_.get(), &p));
                h_ = b200_detail::own(p);
            }
            spline_view_.reset();
            cache_info();
        }
        return *this;
    }
    InterpolationFunction(InterpolationFunction&&) noexcept = default;
    InterpolationFunction& operator=(InterpolationFunction&&) noexcept = default;

    // ---- single point (Interpolation.hpp:132-244); one tiny device launch per call
    template <typename... Coords, typename = std::enable_if_t<sizeof...(Coords) == D &&
                                                              (std::is_arithmetic_v<Coords> && ...)>>
    val_type operator()(Coords... x) const { return (*this)(DimArray<coord_type>{static_cast<coord_type>(x)...}); }
    val_type operator()(DimArray<coord_type> coord) const { return one(coord.data(), nullptr, false); }
    val_type at(DimArray<coord_type> coord) const { return one(coord.data(), nullptr, true); }
    template <typename... Coords, typename = std::enable_if_t<sizeof...(Coords) == D &&
                                                              (std::is_arithmetic_v<Coords> && ...)>>
    val_type at(Coords... x) const { return at(DimArray<coord_type>{static_cast<coord_type>(x)...}); }

    val_type derivative(DimArray<coord_type> coord, DimArray<size_type> derivatives) const {
        const auto dv = to_int(derivatives);
        return one(coord.data(), dv.data(), false);
    }
    template <typename... Args, typename = std::enable_if_t<sizeof...(Args) == D && (std::is_integral_v<Args> && ...)>>
    val_type derivative(DimArray<coord_type> coord, Args... deri) const {
        return derivative(coord, DimArray<size_type>{static_cast<size_type>(deri)...});
    }
    template <typename... P, typename = std::enable_if_t<sizeof...(P) == D>,
              typename = std::void_t<decltype(std::declval<P>().first)...>>
    val_type derivative(P... coord_order) const {
        return derivative(DimArray<coord_type>{static_cast<coord_type>(coord_order.first)...},
                          DimArray<size_type>{static_cast<size_type>(coord_order.second)...});
    }
    val_type derivative_at(DimArray<coord_type> coord, DimArray<size_type> derivatives) const {
        const auto dv = to_int(derivatives);
        return one(coord.data(), dv.data(), true);
    }
    template <typename... Args, typename = std::enable_if_t<sizeof...(Args) == D && (std::is_integral_v<Args> && ...)>>
    val_type derivative_at(DimArray<coord_type> coord, Args... deri) const {
        return derivative_at(coord, DimArray<size_type>{static_cast<size_type>(deri)...});
    }
    template <typename... P, typename = std::enable_if_t<sizeof...(P) == D>,
              typename = std::void_t<decltype(std::declval<P>().first)...>>
    val_type derivative_at(P... coord_order) const {
        return derivative_at(DimArray<coord_type>{static_cast<coord_type>(coord_order.first)...},
                             DimArray<size_type>{static_cast<size_type>(coord_order.second)...});
    }

    // ---- batched entry points (new).  Host pointers: synchronous.  Device pointers: enqueued on
    // `stream` (a cudaStream_t), no synchronisation.
    void evaluate(const coord_type* points, size_type q, val_type* out) const { many(points, q, nullptr, out); }
    void evaluate(const std::vector<DimArray<coord_type>>& points, std::vector<val_type>& out) const {
        out.resize(points.size());
        evaluate(points.empty() ? nullptr : points.front().data(), points.size(), out.data());
    }
    void evaluate(const coord_type* points, size_type q, DimArray<size_type> derivatives, val_type* out) const {
        const auto dv = to_int(derivatives);
        many(points, q, dv.data(), out);
    }
    // out[q][1 + D] = value, d/dx0, ..., d/dx(D-1)
    void evaluate_value_grad(const coord_type* points, size_type q, val_type* out) const {
        if constexpr (b200_detail::direct_v<T, U>) {
            b200_detail::check(bspl_evaluate_value_grad(need(), 0, points, static_cast<int64_t>(q), out, 0, nullptr));
        } else {
            std::vector<coord_type> tmp(q * (1 + D));
            for (size_type k = 0; k < components; ++k) {
                b200_detail::check(bspl_evaluate_value_grad(need(), static_cast<int64_t>(k), points, static_cast<int64_t>(q),
                                                            tmp.data(), 0, nullptr));
                b200_detail::merge_component<T, U>(tmp.data(), q * (1 + D), k, out);
            }
        }
    }
    // device pointers: scalar functions; for vector-valued T evaluate field k with the C ABI (handle())
    void evaluate_device(const coord_type* d_points, size_type q, val_type* d_out, void* stream = nullptr) const {
        static_assert(b200_detail::direct_v<T, U>, "device-pointer evaluation: T must be the coordinate type");
        b200_detail::check(bspl_evaluate(need(), 0, d_points, static_cast<int64_t>(q), nullptr, d_out, 1, stream));
    }
    void evaluate_value_grad_device(const coord_type* d_points, size_type q, val_type* d_out,
                                    void* stream = nullptr) const {
        static_assert(b200_detail::direct_v<T, U>, "device-pointer evaluation: T must be the coordinate type");
        b200_detail::check(bspl_evaluate_value_grad(need(), 0, d_points, static_cast<int64_t>(q), d_out, 1, stream));
    }

    // eval_proxy (Interpolation.hpp:493-506): one point, or a batch of q points [q][D]
    EvalProxy<T, D, O, U> eval_proxy(DimArray<coord_type> coord) const {
        return EvalProxy<T, D, O, U>(need(), coord.data(), 1);
    }
    EvalProxy<T, D, O, U> eval_proxy(const coord_type* points, size_type q) const {
        return EvalProxy<T, D, O, U>(need(), points, q);
    }

    // ---- properties (Interpolation.hpp:248-267)
    bool periodicity(size_type d) const { return periodic_[d]; }
    bool uniform(size_type d) const { return uniform_[d]; }
    const std::pair<coord_type, coord_type>& range(size_type d) const { return range_[d]; }
    static constexpr size_type get_order() { return order; }
    // plain control points of the spline (spline().control_points() without INTP_CELL_LAYOUT)
    Mesh<T, D> control_points() const {
        typename MeshDimension<D>::index_type ext{};
        for (size_type d = 0; d < D; ++d) ext[d] = n_[d];
        Mesh<T, D> m{MeshDimension<D>(ext)};
        if constexpr (b200_detail::direct_v<T, U>) {
            b200_detail::check(bspl_function_control_points(need(), 0, m.data()));
        } else {
            std::vector<coord_type> tmp(m.size());
            for (size_type k = 0; k < components; ++k) {
                b200_detail::check(bspl_function_control_points(need(), static_cast<int64_t>(k), tmp.data()));
                b200_detail::merge_component<T, U>(tmp.data(), m.size(), k, const_cast<T*>(m.data()));
            }
        }
        return m;
    }
    std::vector<coord_type> knots(size_type d) const {
        std::vector<double> k(static_cast<std::size_t>(n_knots_[d]));
        b200_detail::check(bspl_function_knots(need(), static_cast<int>(d), k.data(), static_cast<int64_t>(k.size())));
        return std::vector<coord_type>(k.begin(), k.end());
    }
    const bspl_function* handle() const { return h_.get(); }
    // spline() (Interpolation.hpp:265): the underlying B-spline -- knots, plain control points and the
    // BSpline evaluation interface -- as a view sharing this function's device storage (scalar T).
    // The reference is valid until this function is assigned to or re-interpolated.
    using spline_type = BSpline<T, D, O, U>;
    template <typename S = spline_type>
    const S& spline() const {
        static_assert(b200_detail::direct_v<T, U>, "spline(): T must be the coordinate type");
        if (!spline_view_) {
            DimArray<std::vector<coord_type>> kn;
            for (size_type d = 0; d < D; ++d) kn[d] = knots(d);
            auto sp = std::make_shared<S>(periodic_, control_points(), std::move(kn), range_, h_);
            spline_view_ = sp;
        }
        return *static_cast<const S*>(spline_view_.get());
    }

   private:
    explicit InterpolationFunction(b200_detail::FnHandle h) : h_(std::move(h)) { cache_info(); }
    const bspl_function* need() const {
        if (!h_) throw std::runtime_error("empty InterpolationFunction (populate it with a template's interpolate())");
        return h_.get();
    }
    // one point: component k is field k
    val_type one(const coord_type* c, const int* dv, bool checked) const {
        coord_type buf[components];
        for (size_type k = 0; k < components; ++k) {
            if (checked)
                b200_detail::check(bspl_evaluate_at(need(), static_cast<int64_t>(k), c, 1, dv, &buf[k], nullptr));
            else
                b200_detail::check(bspl_evaluate(need(), static_cast<int64_t>(k), c, 1, dv, &buf[k], 0, nullptr));
        }
        return b200_detail::from_components<T, U>(buf);
    }
    void many(const coord_type* points, size_type q, const int* dv, val_type* out) const {
        if constexpr (b200_detail::direct_v<T, U>) {
            b200_detail::check(bspl_evaluate(need(), 0, points, static_cast<int64_t>(q), dv, out, 0, nullptr));
        } else {
            std::vector<coord_type> tmp(q);
            for (size_type k = 0; k < components; ++k) {
                b200_detail::check(bspl_evaluate(need(), static_cast<int64_t>(k), points, static_cast<int64_t>(q), dv,
                                                 tmp.data(), 0, nullptr));
                b200_detail::merge_component<T, U>(tmp.data(), q, k, out);
            }
        }
    }
    static std::array<int, D> to_int(const DimArray<size_type>& a) {
        std::array<int, D> r{};
        for (size_type d = 0; d < D; ++d) r[d] = static_cast<int>(a[d]);
        return r;
    }
    void cache_info() {
        if (!h_) return;
        int per[BSPL_MAX_DIM], uni[BSPL_MAX_DIM];
        int64_t n[BSPL_MAX_DIM], nk[BSPL_MAX_DIM];
        double lo[BSPL_MAX_DIM], hi[BSPL_MAX_DIM];
        b200_detail::check(bspl_function_info(h_.get(), nullptr, nullptr, nullptr, nullptr, n, per, uni, nk, lo, hi));
        for (size_type d = 0; d < D; ++d) {
            periodic_[d] = per[d] != 0;
            uniform_[d] = uni[d] != 0;
            n_[d] = static_cast<size_type>(n[d]);
            n_knots_[d] = static_cast<size_type>(nk[d]);
            range_[d] = {static_cast<coord_type>(lo[d]), static_cast<coord_type>(hi[d])};
        }
    }

    b200_detail::FnHandle h_;
    mutable std::shared_ptr<const void> spline_view_;  // BSpline<T,D,O,U>, built by spline() on demand
    DimArray<bool> periodic_{};
    DimArray<bool> uniform_{};
    DimArray<size_type> n_{};
    DimArray<size_type> n_knots_{};
    DimArray<std::pair<coord_type, coord_type>> range_{};
};

// ---------------------------------------------------------------- many fields on one mesh (new)
// What InterpolationFunctionTemplate::interpolate_fields returns: F scalar fields interpolated on the
// template's mesh in one batched solve, held as ONE device object, so that a query set is located
// once and applied to every field (the reference's template + eval_proxy use case,
// InterpolationTemplate.hpp:118-176, as one call).
template <typename T, std::size_t D, std::size_t O, typename U = double>
class InterpolationFieldSet {
    static_assert(std::is_same_v<T, U>, "field sets hold scalar fields of the coordinate type");

   public:
    using size_type = std::size_t;
    InterpolationFieldSet() = default;
    size_type size() const { return n_fields_; }
    // out[q] = field k at points[q][D]
    void evaluate(size_type k, const U* points, size_type q, T* out) const {
        b200_detail::check(bspl_evaluate(need(), static_cast<int64_t>(k), points, static_cast<int64_t>(q), nullptr, out, 0, nullptr));
    }
    // out[size()][q]: every field at the same points (host pointers)
    void evaluate_all(const U* points, size_type q, T* out) const {
        b200_detail::check(bspl_evaluate_fields(need(), points, static_cast<int64_t>(q), out, 0, nullptr));
    }
    // the same on device pointers, enqueued on `stream`
    void evaluate_all_device(const U* d_points, size_type q, T* d_out, void* stream = nullptr) const {
        b200_detail::check(bspl_evaluate_fields(need(), d_points, static_cast<int64_t>(q), d_out, 1, stream));
    }
    // out[q][1 + D] = value and gradient of field k
    void evaluate_value_grad(size_type k, const U* points, size_type q, T* out) const {
        b200_detail::check(bspl_evaluate_value_grad(need(), static_cast<int64_t>(k), points, static_cast<int64_t>(q), out, 0, nullptr));
    }
    const bspl_function* handle() const { return h_.get(); }

   private:
    friend class InterpolationFunctionTemplate<T, D, O, U>;
    InterpolationFieldSet(b200_detail::FnHandle h, size_type n) : h_(std::move(h)), n_fields_(n) {}
    const bspl_function* need() const {
        if (!h_) throw std::runtime_error("empty InterpolationFieldSet");
        return h_.get();
    }
    b200_detail::FnHandle h_;
    size_type n_fields_ = 0;
};

// ---------------------------------------------------------------- InterpolationFunctionTemplate
template <typename T, std::size_t D, std::size_t O, typename U = double>
class InterpolationFunctionTemplate {
   public:
    using function_type = InterpolationFunction<T, D, O, U>;
    using size_type = std::size_t;
    using coord_type = U;
    using val_type = T;
    static constexpr size_type dim = D;
    static constexpr size_type order = O;
    template <typename V> using DimArray = std::array<V, D>;
    using MeshDim = MeshDimension<D>;

    // (periodicity, mesh dimension, ranges...)  InterpolationTemplate.hpp:60-79
    template <typename... Ts, typename = std::enable_if_t<sizeof...(Ts) == D>>
    InterpolationFunctionTemplate(DimArray<bool> periodicity, MeshDim mesh_dimension, std::pair<Ts, Ts>... x_ranges)
        : mesh_dimension_(mesh_dimension), periodicity_(periodicity) {
        const std::array<b200_detail::AxisSpec, D> ax{b200_detail::make_axis(x_ranges)...};
        int64_t n[D];
        int per[D];
        double lo[D], hi[D];
        const double* coords[D];
        for (size_type d = 0; d < D; ++d) {
            // data points the library solves for: the dummy sample of a periodic axis is not one of them
            const size_type dummy = (b200_detail::kDummyPoint && periodicity[d]) ? 1 : 0;
            if (mesh_dimension.dim_size(d) <= dummy) throw std::runtime_error("empty axis at dimension " + std::to_string(d));
            n[d] = static_cast<int64_t>(mesh_dimension.dim_size(d) - dummy);
            per[d] = periodicity[d] ? 1 : 0;
            lo[d] = ax[d].lo;
            hi[d] = ax[d].hi;
            coords[d] = ax[d].coords.empty() ? nullptr : ax[d].coords.data();
            // abscissae: one per sample, plus the period's end when that sample is implicit
            // (Interpolation.hpp:379-391)
            if (coords[d] && ax[d].coords.size() != static_cast<size_type>(n[d]) + (periodicity[d] ? 1 : 0))
                throw std::runtime_error("Inconsistency between knot number and interpolated value number at dimension " +
                                         std::to_string(d));
        }
        bspl_template* t = nullptr;
        b200_detail::check(bspl_template_create(b200_detail::dtype_of<U>(), static_cast<int>(D), static_cast<int>(O), n, per,
                                                lo, hi, coords, b200_detail::default_device(), &t));
        h_.reset(t);
    }
    // 1-D forms  InterpolationTemplate.hpp:88-102
    template <typename C1, typename C2, size_type DD = D, typename = std::enable_if_t<DD == 1>>
    InterpolationFunctionTemplate(bool periodicity, size_type f_length, std::pair<C1, C2> x_range)
        : InterpolationFunctionTemplate(DimArray<bool>{periodicity}, MeshDim{f_length},
                                        std::pair<std::common_type_t<C1, C2>, std::common_type_t<C1, C2>>(x_range)) {}
    template <typename C1, typename C2, size_type DD = D, typename = std::enable_if_t<DD == 1>>
    InterpolationFunctionTemplate(size_type f_length, std::pair<C1, C2> x_range)
        : InterpolationFunctionTemplate(false, f_length, x_range) {}
    // all axes non-periodic  InterpolationTemplate.hpp:111-116
    template <typename... Ts, typename = std::enable_if_t<sizeof...(Ts) == D>>
    InterpolationFunctionTemplate(MeshDim mesh_dimension, std::pair<Ts, Ts>... x_ranges)
        : InterpolationFunctionTemplate(DimArray<bool>{}, mesh_dimension, x_ranges...) {}

    // interpolate(mesh)  InterpolationTemplate.hpp:118-133
    function_type interpolate(const Mesh<T, D>& f_mesh) const {
        check_mesh(f_mesh);
        bspl_function* f = nullptr;
        constexpr size_type K = b200_detail::components_of<T, U>::value;
        std::vector<T> kept;
        const T* data = samples(f_mesh, kept);
        [[maybe_unused]] const size_type m = kept.empty() ? f_mesh.size() : kept.size();
        if constexpr (b200_detail::direct_v<T, U>) {
            b200_detail::check(bspl_template_interpolate(h_.get(), data, 1, 0, nullptr, &f));
        } else {  // K fields, one per component
            const std::vector<U> fields = b200_detail::split_components<T, U>(data, m);
            b200_detail::check(bspl_template_interpolate(h_.get(), fields.data(), static_cast<int64_t>(K), 0, nullptr, &f));
        }
        return function_type(b200_detail::own(f));
    }
    template <typename It, size_type DD = D, typename = std::enable_if_t<DD == 1>>
    function_type interpolate(std::pair<It, It> f_range) const {
        return interpolate(Mesh<T, 1>(f_range));
    }
    // interpolate(function&, mesh)  InterpolationTemplate.hpp:136-143
    void interpolate(function_type& interp, const Mesh<T, D>& f_mesh) const {
        check_mesh(f_mesh);
        if (!interp.h_) { interp = interpolate(f_mesh); return; }
        constexpr size_type K = b200_detail::components_of<T, U>::value;
        std::vector<T> kept;
        const T* data = samples(f_mesh, kept);
        [[maybe_unused]] const size_type m = kept.empty() ? f_mesh.size() : kept.size();
        if constexpr (b200_detail::direct_v<T, U>) {
            b200_detail::check(bspl_template_interpolate_into(h_.get(), interp.h_.get(), data, 1, 0, nullptr));
        } else {
            const std::vector<U> fields = b200_detail::split_components<T, U>(data, m);
            b200_detail::check(bspl_template_interpolate_into(h_.get(), interp.h_.get(), fields.data(),
                                                              static_cast<int64_t>(K), 0, nullptr));
        }
        interp.spline_view_.reset();
        interp.cache_info();
    }
    // device-resident mesh, row-major, enqueued on `stream`.  Its shape is that of the data points:
    // a device mesh never carries the dummy sample of a periodic axis.
    function_type interpolate_device(const T* d_mesh, void* stream = nullptr) const {
        static_assert(b200_detail::direct_v<T, U>, "device meshes: T must be the coordinate type (pass K fields through the C ABI)");
        bspl_function* f = nullptr;
        b200_detail::check(bspl_template_interpolate(h_.get(), d_mesh, 1, 1, stream, &f));
        return function_type(b200_detail::own(f));
    }
    // batched interpolate (new): n_fields meshes of this template's data shape stored back to back,
    // [n_fields][n0]...[n_{D-1}] (no dummy samples), solved in one pass per axis
    InterpolationFieldSet<T, D, O, U> interpolate_fields(const T* fields, size_type n_fields) const {
        bspl_function* f = nullptr;
        b200_detail::check(bspl_template_interpolate(h_.get(), fields, static_cast<int64_t>(n_fields), 0, nullptr, &f));
        return InterpolationFieldSet<T, D, O, U>(b200_detail::own(f), n_fields);
    }
    InterpolationFieldSet<T, D, O, U> interpolate_fields_device(const T* d_fields, size_type n_fields,
                                                                void* stream = nullptr) const {
        bspl_function* f = nullptr;
        b200_detail::check(bspl_template_interpolate(h_.get(), d_fields, static_cast<int64_t>(n_fields), 1, stream, &f));
        return InterpolationFieldSet<T, D, O, U>(b200_detail::own(f), n_fields);
    }
    // eval_proxy (InterpolationTemplate.hpp:145-176): everything that depends on the point only,
    // done before the fields exist; proxy(function) then evaluates any function of this template.
    using eval_proxy_t = EvalProxy<T, D, O, U>;
    eval_proxy_t eval_proxy(DimArray<coord_type> coord) const { return eval_proxy_t(h_.get(), coord.data(), 1); }
    template <typename... Coords, typename = std::enable_if_t<sizeof...(Coords) == D &&
                                                              (std::is_arithmetic_v<Coords> && ...)>>
    eval_proxy_t eval_proxy(Coords... x) const {
        return eval_proxy(DimArray<coord_type>{static_cast<coord_type>(x)...});
    }
    // batched: q points [q][D]
    eval_proxy_t eval_proxy(const coord_type* points, size_type q) const { return eval_proxy_t(h_.get(), points, q); }

    const MeshDim& mesh_dimension() const { return mesh_dimension_; }
    const bspl_template* handle() const { return h_.get(); }

   private:
    void check_mesh(const Mesh<T, D>& m) const {
        for (size_type d = 0; d < D; ++d)
            if (m.dim_size(d) != mesh_dimension_.dim_size(d)) throw std::runtime_error("mesh shape differs from the template's");
    }
    // the samples handed to the library: the mesh itself, or (dummy-point mode, periodic axes) a copy
    // without the closing sample of every periodic axis, kept alive in `kept`
    const T* samples(const Mesh<T, D>& f_mesh, std::vector<T>& kept) const {
        if constexpr (b200_detail::kDummyPoint) {
            for (size_type d = 0; d < D; ++d)
                if (periodicity_[d]) {
                    kept = b200_detail::strip_dummy<T, D>(f_mesh.data(), mesh_dimension_, periodicity_);
                    return kept.data();
                }
        }
        return f_mesh.data();
    }
    MeshDim mesh_dimension_;
    DimArray<bool> periodicity_{};
    b200_detail::TmHandle h_;
};

// ---------------------------------------------------------------- 1-D conveniences
template <std::size_t O = 3, typename T = double, typename U = double>
class InterpolationFunction1D : public InterpolationFunction<T, 1, O, U> {
    using base = InterpolationFunction<T, 1, O, U>;

   public:
    // default x range [0, N-1]; [0, N] for a periodic axis whose closing sample is implicit
    // (INTP_PERIODIC_NO_DUMMY_POINT)  Interpolation.hpp:518-533
    template <typename It>
    InterpolationFunction1D(std::pair<It, It> f_range, bool periodicity = false)
        : InterpolationFunction1D(
              std::make_pair(U{}, static_cast<U>(std::distance(f_range.first, f_range.second) -
                                                 ((periodicity && !b200_detail::kDummyPoint) ? 0 : 1))),
              f_range, periodicity) {}
    template <typename C1, typename C2, typename It>
    InterpolationFunction1D(std::pair<C1, C2> x_range, std::pair<It, It> f_range, bool periodicity = false)
        : base(periodicity, f_range, x_range) {}
};

template <std::size_t O, typename T = double, typename U = double>
class InterpolationFunctionTemplate1D : public InterpolationFunctionTemplate<T, 1, O, U> {
    using base = InterpolationFunctionTemplate<T, 1, O, U>;

   public:
    InterpolationFunctionTemplate1D(typename base::size_type f_length, bool periodicity = false)
        : InterpolationFunctionTemplate1D(std::make_pair(U{}, static_cast<U>(f_length - 1)), f_length, periodicity) {}
    template <typename C1, typename C2>
    InterpolationFunctionTemplate1D(std::pair<C1, C2> x_range, typename base::size_type f_length, bool periodicity = false)
        : base(periodicity, f_length, x_range) {}
};

}  // namespace intp

#endif  // INTP_B200_INTERPOLATION_HPP
