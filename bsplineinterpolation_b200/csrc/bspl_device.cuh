// Device-side building blocks shared by the evaluation kernels: per-axis
// parameters, uniform-knot reconstruction, knot-span location and Cox-de Boor
// weights.  Everything that decides a span is written with explicit
// round-to-nearest intrinsics (no FMA contraction) so that it reproduces the
// reference's x86 arithmetic bit for bit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bspl {

constexpr int kMaxDim = 4;
constexpr int kMaxOrder = 7;

// Per-axis description handed to kernels by value.
template <typename R>
struct AxisParams {
    const R* t;      // knot array on the device, or nullptr on uniform axes
    R lo, hi;        // user range [a, b]                       (uniform axes)
    R dx;            // (b - a) / (n' - 1)                      Interpolation.hpp:335
    R half_extra;    // 0.5 * extra, extra = knots - samples    Interpolation.hpp:338
    R inv_dx;
    R first, second; // range(): wrap interval of periodic axes BSpline.hpp:224-226
    int n;           // control points
    int K;           // knots
    int periodic;
    int unit_ok;     // uniform axis whose knots are small multiples of dx: weights may use unit-spaced knots
    long long stride;  // element stride of this axis in the padded coefficient array
};

template <typename R> struct Arith;
template <> struct Arith<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double fmodr(double a, double b) { return fmod(a, b); }
    static __device__ __forceinline__ double floorr(double a) { return floor(a); }
};
template <> struct Arith<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float fmodr(float a, float b) { return fmodf(a, b); }
    static __device__ __forceinline__ float floorr(float a) { return floorf(a); }
};

// Knot i of the axis.  Uniform axes rebuild the reference's value
//   t[i] = a + (double(i) - 0.5*extra) * dx        (Interpolation.hpp:343-354)
// with the clamped end knots of non-periodic axes set exactly (:340, :356-358).
template <typename R, int O>
__device__ __forceinline__ R knot_at(const AxisParams<R>& a, int i) {
    if (a.t != nullptr) return a.t[i];
    if (!a.periodic) {
        if (i <= O) return a.lo;
        if (i >= a.K - O - 1) return a.hi;
    }
    using A = Arith<R>;
    return A::add(a.lo, A::mul(A::sub(static_cast<R>(i), a.half_extra), a.dx));
}

// get_knot_iter (BSpline.hpp:125-157).  Wraps x into [first, second) on periodic
// axes (x is updated, as in the reference) and returns
//   span = O + #{ i in [O+1, K-O-2] : t[i] <= x },
// which is what both the hint-accept branch and the upper_bound branch of the
// reference produce for strictly increasing interior knots.
template <typename R, int O>
__device__ __forceinline__ int locate(const AxisParams<R>& a, R& x) {
    using A = Arith<R>;
    if (a.periodic) {
        const R period = A::sub(a.second, a.first);
        const R x0 = x;
        x = A::add(A::add(a.first, A::fmodr(A::sub(x0, a.first), period)),
                   x0 < a.first ? period : R(0));
    }
    const int lo_i = O, hi_i = a.K - O - 2;
    if (a.t == nullptr) {
        // one multiply for the guess, then exact fix-up against the true knots
        R g = A::floorr((x - a.lo) * a.inv_dx + a.half_extra);
        g = fmax(g, static_cast<R>(lo_i));   // also maps NaN to lo_i
        g = fmin(g, static_cast<R>(hi_i));
        int s = static_cast<int>(g);
        while (s < hi_i && knot_at<R, O>(a, s + 1) <= x) ++s;
        while (s > lo_i && knot_at<R, O>(a, s) > x) --s;
        return s;
    }
    // general knots: first index in [O+1, hi_i+1) with t > x, minus one
    int lo = O + 1, hi = hi_i + 1;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (!(x < a.t[mid])) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

// locate() with a short cut for uniform axes: when the guessed span and its neighbours lie among
// the formula knots (away from the clamped ends) the answer is decided by at most four knot
// values computed without the clamp logic; anything else falls back to locate().  Same knots,
// same comparisons, hence the same span.  x is NOT updated (periodic wrap stays local).
template <typename R, int O>
__device__ __forceinline__ int locate_quick(const AxisParams<R>& a, R x) {
    using A = Arith<R>;
    if (a.t == nullptr) {
        R xw = x;
        if (a.periodic) {
            const R period = A::sub(a.second, a.first);
            xw = A::add(A::add(a.first, A::fmodr(A::sub(x, a.first), period)), x < a.first ? period : R(0));
        }
        const R g = A::floorr((xw - a.lo) * a.inv_dx + a.half_extra);
        const R g_lo = static_cast<R>(a.periodic ? O + 1 : O + 2), g_hi = static_cast<R>(a.K - O - 4);
        if (g >= g_lo && g <= g_hi) {  // knots g-1 .. g+2 are formula knots and g-1 .. g+1 valid spans
            const R t0 = A::add(a.lo, A::mul(A::sub(g, a.half_extra), a.dx));
            const R t1 = A::add(a.lo, A::mul(A::sub(g + R(1), a.half_extra), a.dx));
            const int s = static_cast<int>(g);
            if (t0 <= xw) {
                if (xw < t1) return s;
                const R t2 = A::add(a.lo, A::mul(A::sub(g + R(2), a.half_extra), a.dx));
                if (xw < t2) return s + 1;
            } else {
                const R tm = A::add(a.lo, A::mul(A::sub(g - R(1), a.half_extra), a.dx));
                if (tm <= xw) return s - 1;
            }
        }
    }
    R xx = x;
    return locate<R, O>(a, xx);
}

// Local knot window tk[m] = t[span - O + 1 + m], m = 0 .. 2O-1.
template <typename R, int O>
__device__ __forceinline__ void load_knot_window(const AxisParams<R>& a, int span, R* tk) {
#pragma unroll
    for (int m = 0; m < 2 * O; ++m) tk[m] = knot_at<R, O>(a, span - O + 1 + m);
}

// Reciprocal to within ~1 ulp without the IEEE division sequence: hardware seed
// (MUFU.RCP64H / MUFU.RCP) refined by Newton steps.  Used only for basis weights
// (tolerance 1e-12), never for span selection.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ float fast_rcp(float d) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    const float e = fmaf(-d, r, 1.0f);
    return fmaf(r, e, r);
}

// One level of the Cox-de Boor triangle (BSpline.hpp:91-109) in place: on entry b
// holds the order i-1 functions right-aligned in b[0..O]; inv[j] = 1 / (t[span+j] -
// t[span+j-i]), j = 1..i, are the level's distinct denominators.
template <typename R, int O>
__device__ __forceinline__ void basis_level(const R* tk, R x, int i, const R* inv, R* b) {
    const int ib = O - i;
#pragma unroll
    for (int j = 0; j <= O; ++j) {
        if (j <= i) {
            const int l = O - 1 - (i - j);  // local index of t[span-(i-j)]
            const int r = O + j;            // local index of t[span+j+1]
            R left = R(0), right = R(0);
            if (j != 0) left = b[ib + j] * (x - tk[l]) * inv[j];
            if (ib + j != O) right = b[ib + j + 1] * (tk[r] - x) * inv[j + 1];
            b[ib + j] = left + right;
        }
    }
}

template <typename R, int O>
__device__ __forceinline__ void level_reciprocals(const R* tk, int i, R* inv) {
#pragma unroll
    for (int j = 1; j <= O; ++j)
        if (j <= i) inv[j] = fast_rcp(tk[O - 1 + j] - tk[O - 1 + j - i]);
}

// ---- fast path for uniform axes away from the clamped end knots -------------------
// Valid when every knot the query touches (span-O+1 .. span+O, and span+1 for the
// fix-up) is a formula knot, i.e. O + 1 <= index <= K-O-2 on non-periodic axes (always
// on periodic ones).  The knot VALUES are the reference's, bit for bit -- (double(i)-h)
// is exact, so stepping the index in floating point changes nothing -- which keeps the
// span selection exact; only the divisions of the weights are replaced by a
// Newton-refined reciprocal seeded with 1/(i dx).
template <typename R>
__device__ __forceinline__ R uniform_knot(const AxisParams<R>& a, R fi) {
    using A = Arith<R>;
    return A::add(a.lo, A::mul(A::sub(fi, a.half_extra), a.dx));
}

// Returns span; fs = R(span); x wrapped on periodic axes.  lo_i/hi_i as in locate().
template <typename R, int O>
__device__ __forceinline__ int locate_uniform_interior(const AxisParams<R>& a, R& x, R& fs) {
    using A = Arith<R>;
    if (a.periodic) {
        const R period = A::sub(a.second, a.first);
        const R x0 = x;
        x = A::add(A::add(a.first, A::fmodr(A::sub(x0, a.first), period)),
                   x0 < a.first ? period : R(0));
    }
    const int lo_i = O, hi_i = a.K - O - 2;
    R g = A::floorr((x - a.lo) * a.inv_dx + a.half_extra);
    g = fmin(fmax(g, static_cast<R>(lo_i)), static_cast<R>(hi_i));
    int s = static_cast<int>(g);
    fs = g;
    while (s < hi_i && uniform_knot<R>(a, fs + R(1)) <= x) { ++s; fs += R(1); }
    while (s > lo_i && uniform_knot<R>(a, fs) > x) { --s; fs -= R(1); }
    return s;
}

template <typename R, int O>
__device__ __forceinline__ void uniform_knot_window(const AxisParams<R>& a, R fs, R* tk) {
#pragma unroll
    for (int m = 0; m < 2 * O; ++m) tk[m] = uniform_knot<R>(a, fs + R(m - O + 1));
}

template <typename R, int O>
__device__ __forceinline__ void level_reciprocals_uniform(const R* tk, int i, R inv_dx, R* inv) {
    const R seed = inv_dx * (R(1) / R(i));  // i is a compile-time constant after unrolling
#pragma unroll
    for (int j = 1; j <= O; ++j)
        if (j <= i) {
            const R den = tk[O - 1 + j] - tk[O - 1 + j - i];
            const R e = fma(-den, seed, R(1));
            inv[j] = fma(seed, e, seed);
        }
}

// basis_and_deriv with the uniform-seeded reciprocals.
template <typename R, int O, bool GRAD>
__device__ __forceinline__ void basis_uniform(const R* tk, R x, R inv_dx, R* w, R* dw) {
#pragma unroll
    for (int i = 0; i <= O; ++i) { w[i] = R(0); if (GRAD) dw[i] = R(0); }
    w[O] = R(1);
    if (O == 0) return;
    R inv[O + 2];
#pragma unroll
    for (int i = 1; i < O; ++i) {
        level_reciprocals_uniform<R, O>(tk, i, inv_dx, inv);
        basis_level<R, O>(tk, x, i, inv, w);
    }
    level_reciprocals_uniform<R, O>(tk, O, inv_dx, inv);
    if (GRAD) {
#pragma unroll
        for (int j = 0; j <= O; ++j) {
            R v = R(0);
            if (j >= 1) v = w[j] * inv[j];
            if (j + 1 <= O) v -= w[j + 1] * inv[j + 1];
            dw[j] = R(O) * v;
        }
    }
    basis_level<R, O>(tk, x, O, inv, w);
}

// Cox-de Boor on the unit-spaced knots of a uniform interior cell.  With u = (x - t[span]) / dx
// the local knots are the integers, every denominator of level j is exactly j, and the triangle
// needs no knot values and no reciprocals: N[r] <- saved + (r+1-u) * N[r]/j, saved <- (u+j-r-1) * N[r]/j.
// w[r] weights control point span - O + r; dw = d w / d x = (N'[r-1] - N'[r]) / dx from the
// level-(O-1) values N'.  Agrees with the knot-based triangle to a few ulp * (|t| / dx).
template <typename R, int O, bool GRAD>
__device__ __forceinline__ void basis_unit(R u, R inv_dx, R* w, R* dw) {
#pragma unroll
    for (int i = 0; i <= O; ++i) { w[i] = R(0); if (GRAD) dw[i] = R(0); }
    // N[0..j] after level j, kept left-aligned while building, moved right-aligned at the end
    R N[O + 1];
    N[0] = R(1);
#pragma unroll
    for (int j = 1; j <= O; ++j) {
        if (GRAD && j == O) {
            // N holds the level-(O-1) values N'[0..O-1] of functions span-O+1 .. span
#pragma unroll
            for (int r = 0; r <= O; ++r) {
                R v = R(0);
                if (r >= 1) v = N[r - 1];
                if (r <= O - 1) v -= N[r];
                dw[r] = v * inv_dx;
            }
        }
        const R inv_j = R(1) / R(j);  // compile-time constant after unrolling
        R saved = R(0);
#pragma unroll
        for (int r = 0; r < j; ++r) {
            const R temp = N[r] * inv_j;
            N[r] = fma(R(r + 1) - u, temp, saved);
            saved = (u + R(j - r - 1)) * temp;
        }
        N[j] = saved;
    }
#pragma unroll
    for (int r = 0; r <= O; ++r) w[r] = N[r];
}

// base_spline_value (BSpline.hpp:83-111): Cox-de Boor triangle of order `so`
// (so <= O), result right-aligned in b[0..O].
template <typename R, int O>
__device__ __forceinline__ void basis_funs(const R* tk, R x, int so, R* b) {
#pragma unroll
    for (int i = 0; i <= O; ++i) b[i] = R(0);
    b[O] = R(1);
    R inv[O + 2];
#pragma unroll
    for (int i = 1; i <= O; ++i) {
        if (i <= so) {
            level_reciprocals<R, O>(tk, i, inv);
            basis_level<R, O>(tk, x, i, inv, b);
        }
    }
}

// Value weights w and first-derivative weights dw of one axis from a single pass
// over the triangle: the order O-1 functions are an intermediate of the order O
// ones, and both use the last level's denominators
//   dw[j] = O * ( B^{O-1}[j] / (t[span+j]-t[span+j-O]) - B^{O-1}[j+1] / (t[span+j+1]-t[span+j+1-O]) ).
template <typename R, int O>
__device__ __forceinline__ void basis_and_deriv(const R* tk, R x, R* w, R* dw) {
#pragma unroll
    for (int i = 0; i <= O; ++i) { w[i] = R(0); dw[i] = R(0); }
    w[O] = R(1);
    if (O == 0) return;
    R inv[O + 2];
#pragma unroll
    for (int i = 1; i < O; ++i) {
        level_reciprocals<R, O>(tk, i, inv);
        basis_level<R, O>(tk, x, i, inv, w);
    }
    level_reciprocals<R, O>(tk, O, inv);
#pragma unroll
    for (int j = 0; j <= O; ++j) {
        R v = R(0);
        if (j >= 1) v = w[j] * inv[j];
        if (j + 1 <= O) v -= w[j + 1] * inv[j + 1];
        dw[j] = R(O) * v;
    }
    basis_level<R, O>(tk, x, O, inv, w);
}

// Weights of the k-th derivative along one axis: w such that
//   d^k/dx^k sum_j c[j] B_j(x) = sum_j w[j] c[j].
// The reference differences the local control points k times
// (BSpline.hpp:507-518: c[i] <- m (c[i]-c[i-1]) / (t[i-O+m] - t[i-O]), m = O..O-k+1)
// and dots them with the order O-k basis; applying the transposed stages to
// that basis gives the same number with the control points left untouched.
template <typename R, int O>
__device__ __forceinline__ void deriv_weights(const R* tk, R x, int k, R* w) {
    basis_funs<R, O>(tk, x, O - k, w);
#pragma unroll
    for (int m = 1; m <= O; ++m) {
        if (m > O - k) {
#pragma unroll
            for (int i = 0; i <= O; ++i) {
                if (i >= O - m) {
                    // a_i(m) = m / (t[span+i-O+m] - t[span+i-O]); local idx = global - (span-O+1)
                    R v = R(0);
                    if (i >= O + 1 - m) v = w[i] * (R(m) * fast_rcp(tk[i + m - 1] - tk[i - 1]));
                    if (i + 1 <= O) v -= w[i + 1] * (R(m) * fast_rcp(tk[i + m] - tk[i]));
                    w[i] = v;
                }
            }
        }
    }
}

}  // namespace bspl
