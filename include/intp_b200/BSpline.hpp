// intp_b200/BSpline.hpp -- the reference's BSpline<T, D, O, U> (src/include/BSpline.hpp:24-760) as a
// host handle on a device-resident spline: knot vectors and control points go in through the same
// constructors and load_* calls, every value and derivative comes from the kernels of
// libbspline_b200.so (bspl_function_from_control_points / bspl_evaluate).  Position hints of the
// reference's overloads are accepted and ignored: the device locate is exact without them.
// control_points() returns the plain array (the reference's answer without INTP_CELL_LAYOUT); the
// cell layout is an internal device matter here.
#ifndef INTP_B200_BSPLINE_HPP
#define INTP_B200_BSPLINE_HPP

#include <array>
#include <cstdint>
#include <iostream>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../bspline_b200.h"
#include "Mesh.hpp"
#include "util.hpp"

namespace intp {

template <typename T, std::size_t D, std::size_t O, typename U = double>
class BSpline {
    static_assert(D >= 1 && D <= BSPL_MAX_DIM && O <= BSPL_MAX_ORDER, "dim 1..4, order 0..7 (the reference takes any; these are the instantiated device kernels)");
    static_assert(std::is_same_v<T, U> && (std::is_same_v<T, double> || std::is_same_v<T, float>),
                  "BSpline: T and U must both be double or both be float");

   public:
    using spline_type = BSpline<T, D, O, U>;
    using size_type = std::size_t;
    using val_type = T;
    using knot_type = U;
    static constexpr size_type dim = D;
    static constexpr size_type order = O;
    using KnotContainer = std::vector<knot_type>;
    using ControlPointContainer = Mesh<val_type, D>;
    using control_point_type = ControlPointContainer;
    using knot_const_iterator = typename KnotContainer::const_iterator;
    template <typename V>
    using DimArray = std::array<V, D>;

    explicit BSpline(DimArray<bool> periodicity) : periodicity_(periodicity), ctrl_(size_type{0}) {}
    explicit BSpline() : BSpline(DimArray<bool>{}) {}

    // (periodicity, control points, one (begin, end) pair of knots per axis)  BSpline.hpp:188-210
    template <typename... InputIters, typename = std::enable_if_t<sizeof...(InputIters) == D>>
    BSpline(DimArray<bool> periodicity, ControlPointContainer ctrl_pts, std::pair<InputIters, InputIters>... knot_iter_pairs)
        : periodicity_(periodicity),
          knots_{KnotContainer(knot_iter_pairs.first, knot_iter_pairs.second)...},
          ctrl_(std::move(ctrl_pts)) {
        for (size_type d = 0; d < D; ++d) {
            const size_type want = periodicity_[d] ? 2 * O + 1 : O + 1;
            if (knots_[d].size() < ctrl_.dim_size(d) || knots_[d].size() - ctrl_.dim_size(d) != want)
                throw std::runtime_error("Inconsistency between knot number and control point number at dimension " +
                                         std::to_string(d));
            set_range(d, O + 1);  // BSpline.hpp:200-202
        }
    }
    template <typename... InputIters, typename = std::enable_if_t<sizeof...(InputIters) == D>>
    BSpline(ControlPointContainer ctrl_pts, std::pair<InputIters, InputIters>... knot_iter_pairs)
        : BSpline(DimArray<bool>{}, std::move(ctrl_pts), knot_iter_pairs...) {}

    // A view of an existing device spline (what InterpolationFunction::spline() returns): the host
    // copies of knots and control points describe it, evaluation goes to the shared device storage.
    BSpline(DimArray<bool> periodicity, ControlPointContainer ctrl_pts, DimArray<KnotContainer> knots,
            DimArray<std::pair<knot_type, knot_type>> ranges, std::shared_ptr<bspl_function> device_spline)
        : periodicity_(periodicity), knots_(std::move(knots)), ctrl_(std::move(ctrl_pts)), range_(ranges),
          device_(std::move(device_spline)) {}

    // BSpline.hpp:217-242
    void load_knots(size_type d, KnotContainer knots) {
        knots_[d] = std::move(knots);
        set_range(d, O + (2 - O % 2));  // BSpline.hpp:224-226
        device_.reset();
    }
    void load_ctrlPts(ControlPointContainer control_points) {
        ctrl_ = std::move(control_points);
        device_.reset();
    }

    // ---- value  BSpline.hpp:305-390
    val_type operator()(DimArray<double> coords) const {
        DimArray<knot_type> x;
        for (size_type d = 0; d < D; ++d) x[d] = static_cast<knot_type>(coords[d]);
        return one(x, nullptr);
    }
    val_type operator()(DimArray<std::pair<knot_type, size_type>> coord_with_hints) const {
        DimArray<knot_type> x;
        for (size_type d = 0; d < D; ++d) x[d] = coord_with_hints[d].first;
        return one(x, nullptr);
    }

    // ---- derivative  BSpline.hpp:393-550: (coordinate, derivative order) per axis, optionally with a
    // position hint in the middle
    val_type derivative_at(DimArray<std::pair<knot_type, size_type>> coord_deriOrders) const {
        DimArray<knot_type> x;
        std::array<int, D> dv;
        for (size_type d = 0; d < D; ++d) {
            x[d] = coord_deriOrders[d].first;
            dv[d] = static_cast<int>(coord_deriOrders[d].second);
        }
        return one(x, dv.data());
    }
    val_type derivative_at(DimArray<std::tuple<knot_type, size_type, size_type>> coord_hint_deriOrder) const {
        DimArray<knot_type> x;
        std::array<int, D> dv;
        for (size_type d = 0; d < D; ++d) {
            x[d] = std::get<0>(coord_hint_deriOrder[d]);
            dv[d] = static_cast<int>(std::get<2>(coord_hint_deriOrder[d]));
        }
        return one(x, dv.data());
    }

    // ---- pre_calc_coef (BSpline.hpp:244-297): the point-dependent work done once.  The returned
    // closure evaluates this spline (or a copy sharing its device storage) at the point.
    class Evaluator {
       public:
        val_type operator()(const spline_type& spline) const {
            val_type v{};
            check(bspl_query_plan_evaluate(plan_.get(), spline.handle(), 0, nullptr, 0, &v, 0, nullptr));
            return v;
        }

       private:
        friend class BSpline;
        struct PlanDeleter { void operator()(bspl_query_plan* p) const { bspl_query_plan_destroy(p); } };
        std::shared_ptr<bspl_query_plan> plan_;
    };
    Evaluator pre_calc_coef(DimArray<std::pair<knot_type, size_type>> coord_with_hints) const {
        DimArray<knot_type> x;
        for (size_type d = 0; d < D; ++d) x[d] = coord_with_hints[d].first;
        bspl_query_plan* p = nullptr;
        check(bspl_query_plan_create(device(), x.data(), 1, 0, nullptr, &p));
        Evaluator e;
        e.plan_.reset(p, typename Evaluator::PlanDeleter());
        return e;
    }

    // ---- batched (new): points [q][D] -> out [q]; derivatives == nullptr for values
    void evaluate(const knot_type* points, size_type q, val_type* out, const DimArray<size_type>* derivatives = nullptr) const {
        std::array<int, D> dv{};
        if (derivatives)
            for (size_type d = 0; d < D; ++d) dv[d] = static_cast<int>((*derivatives)[d]);
        check(bspl_evaluate(device(), 0, points, static_cast<int64_t>(q), derivatives ? dv.data() : nullptr, out, 0, nullptr));
    }

    // ---- properties  BSpline.hpp:560-608
    knot_const_iterator knots_begin(size_type d) const { return knots_[d].cbegin(); }
    knot_const_iterator knots_end(size_type d) const { return knots_[d].cend(); }
    const control_point_type& control_points() const { return ctrl_; }
    const std::pair<knot_type, knot_type>& range(size_type d) const { return range_[d]; }
    size_type knots_num(size_type d) const { return knots_[d].size(); }
    bool periodicity(size_type d) const { return periodicity_[d]; }
    constexpr size_type get_order() const { return order; }
    void set_device(int ordinal) { device_ordinal_ = ordinal; device_.reset(); }
    // control points, one row of the last axis per line (BSpline.hpp:611-628, there under INTP_DEBUG)
    void debug_output(std::ostream& os = std::cout) const {
        os << "\n[DEBUG] Control Points (raw data):\n";
        const auto prec = os.precision(17);
        const size_type row = ctrl_.dim_size(D - 1);
        size_type idx = 0;
        for (auto v : ctrl_) {
            if (idx % row == 0) os << "[DEBUG] ";
            os << v << ' ';
            if (++idx % row == 0) os << '\n';
        }
        os << '\n';
        os.precision(prec);
    }
    // the device-resident spline (built on first use)
    const bspl_function* handle() const { return device(); }

   private:
    struct Deleter { void operator()(bspl_function* p) const { bspl_function_destroy(p); } };

    static void check(int rc) {
        if (rc == BSPL_OK) return;
        if (rc == BSPL_ERR_ALLOC) throw std::bad_alloc();
        throw std::runtime_error(bspl_last_error());
    }
    // range = [t[O], t[K - back]]: the reference takes back = O + 1 in its constructor and
    // O + (2 - O % 2) in load_knots (they differ for even orders); the device spline wraps periodic
    // coordinates with the constructor's convention.
    void set_range(size_type d, size_type back) {
        const KnotContainer& t = knots_[d];
        if (t.size() < 2 * O + 2) throw std::runtime_error("too few knots at dimension " + std::to_string(d));
        range_[d] = {t[O], t[t.size() - back]};
    }
    const bspl_function* device() const {
        if (!device_) {
            int64_t n_ctrl[D], n_knots[D];
            int per[D];
            std::array<std::vector<double>, D> kd;
            const double* kp[D];
            for (size_type d = 0; d < D; ++d) {
                if (knots_[d].empty() || ctrl_.size() == 0)
                    throw std::runtime_error("BSpline: knots and control points must be loaded before evaluation");
                n_ctrl[d] = static_cast<int64_t>(ctrl_.dim_size(d));
                n_knots[d] = static_cast<int64_t>(knots_[d].size());
                per[d] = periodicity_[d] ? 1 : 0;
                kd[d].assign(knots_[d].begin(), knots_[d].end());
                kp[d] = kd[d].data();
            }
            bspl_function* f = nullptr;
            check(bspl_function_from_control_points(std::is_same_v<T, double> ? BSPL_F64 : BSPL_F32, static_cast<int>(D),
                                                    static_cast<int>(O), n_ctrl, per, kp, n_knots, ctrl_.data(), 1,
                                                    device_ordinal_, &f));
            device_.reset(f, Deleter());
        }
        return device_.get();
    }
    val_type one(const DimArray<knot_type>& x, const int* dv) const {
        val_type v{};
        check(bspl_evaluate(device(), 0, x.data(), 1, dv, &v, 0, nullptr));
        return v;
    }

    DimArray<bool> periodicity_{};
    DimArray<KnotContainer> knots_{};
    ControlPointContainer ctrl_;
    DimArray<std::pair<knot_type, knot_type>> range_{};
    int device_ordinal_ = 0;
    mutable std::shared_ptr<bspl_function> device_;  // copies share it; any load_* drops it
};

}  // namespace intp

#endif  // INTP_B200_BSPLINE_HPP
