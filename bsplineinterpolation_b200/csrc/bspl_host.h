// Host-side model of one spline axis: knot construction, collocation matrix
// assembly and band LU, all in the scalar type R of the spline so that the
// numbers equal the reference's (which computes them in coord_type).
// Internal to the library; restates
//   create_knot_vector_   Interpolation.hpp:322-362 (uniform), :365-464 (non-uniform)
//   load_knots            BSpline.hpp:217-227
//   build_solver_         InterpolationTemplate.hpp:254-446
//   BandLU::compute_impl  BandLU.hpp:103-118, :159-213
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <new>
#include <stdexcept>
#include <vector>

namespace bspl {

template <typename R>
struct HostAxis {
    int order = 0;
    bool periodic = false;
    bool uniform = true;
    int64_t n = 0;   // data points == control points
    int64_t K = 0;   // knots
    R lo = 0, hi = 0, dx = 0, half_extra = 0;
    R first = 0, second = 0;
    std::vector<R> t;       // explicit knots (non-uniform axes / from-knots splines)
    std::vector<R> coords;  // data abscissae (non-uniform axes)

    R knot(int64_t i) const {
        if (!uniform) return t[static_cast<size_t>(i)];
        if (!periodic) {
            if (i <= order) return lo;
            if (i >= K - order - 1) return hi;
        }
        // volatile-free, but written as three separately rounded operations;
        // host code is built without FMA contraction (see build flags).
        const R a = static_cast<R>(i) - half_extra;
        const R b = a * dx;
        return lo + b;
    }

    void set_uniform(int order_, bool periodic_, int64_t n_, R lo_, R hi_) {
        order = order_; periodic = periodic_; uniform = true; n = n_; lo = lo_; hi = hi_;
        const int64_t samples = n + (periodic ? 1 : 0);  // INTP_PERIODIC_NO_DUMMY_POINT
        if (samples < 2) throw std::invalid_argument("axis needs at least two samples");
        dx = (hi - lo) / static_cast<R>(samples - 1);
        const int64_t extra = periodic ? 2 * order + (1 - order % 2) : order + 1;
        half_extra = R(.5) * static_cast<R>(extra);
        K = samples + extra;
        finish_range(K - order - (2 - order % 2));
    }

    void set_nonuniform(int order_, bool periodic_, int64_t n_, const double* c) {
        order = order_; periodic = periodic_; uniform = false; n = n_;
        const int64_t m = n + (periodic ? 1 : 0);  // abscissae supplied
        const int O = order;
        if (!periodic && O == 0) throw std::invalid_argument("order 0 needs a uniform or periodic axis");
        K = periodic ? m + 2 * O + (1 - O % 2) : m + O + 1;
        t.assign(static_cast<size_t>(K), R(0));
        coords.resize(static_cast<size_t>(m));
        for (int64_t i = 0; i < m; ++i) coords[i] = static_cast<R>(c[i]);
        for (int64_t i = 0; i + 1 < m; ++i)
            if (!(coords[i + 1] > coords[i])) throw std::invalid_argument("coordinates must increase");
        if (periodic) {
            // interior knots are the abscissae (odd O) or their midpoints (even O), then
            // extended by one period on both sides
            for (int64_t i = 1; i < m; ++i)
                t[O + i] = (O % 2 == 0) ? R(.5) * (coords[i - 1] + coords[i]) : coords[i];
            const R period = coords[m - 1] - coords[0];
            for (int64_t i = 0; i <= O; ++i) {
                t[i] = t[m + i - 1] - period;
                t[K - i - 1] = t[K - i - m] + period;
            }
        } else {
            // clamped ends; interior knot i is the mean of O consecutive abscissae
            for (int64_t i = 0; i <= O; ++i) t[i] = coords[0];
            R window = 0;
            for (int64_t i = 1; i < O; ++i) window += coords[i];
            for (int64_t i = O + 1; i < m; ++i) {
                window += coords[i - 1];
                t[i] = window / static_cast<R>(O);
                window -= coords[i - O];
            }
            for (int64_t i = m; i < m + O + 1; ++i) t[i] = coords[m - 1];
        }
        finish_range(K - order - (2 - order % 2));
    }

    void set_from_knots(int order_, bool periodic_, int64_t n_ctrl, const double* k, int64_t nk) {
        order = order_; periodic = periodic_; uniform = false; n = n_ctrl; K = nk;
        if (nk - n_ctrl != (periodic ? 2 * order + 1 : order + 1))
            throw std::invalid_argument("knot count does not match control point count");
        t.resize(static_cast<size_t>(nk));
        for (int64_t i = 0; i < nk; ++i) t[i] = static_cast<R>(k[i]);
        finish_range(K - order - 1);  // BSpline.hpp:200-202
    }

    // first index s in [O, K-O-2] with t[s] <= x < t[s+1] semantics of
    // get_knot_iter(x, hint, last) (BSpline.hpp:125-157), no wrap.
    int64_t span_of(R x, int64_t hint, int64_t last) const {
        if (knot(hint) <= x && knot(hint + 1) > x) return hint;
        int64_t a = order + 1, b = last + 1;
        while (a < b) {
            const int64_t mid = a + (b - a) / 2;
            if (!(x < knot(mid))) a = mid + 1; else b = mid;
        }
        return a - 1;
    }

    // base_spline_value (BSpline.hpp:83-111)
    void basis(int64_t seg, R x, R* b) const {
        const int O = order;
        for (int i = 0; i <= O; ++i) b[i] = 0;
        b[O] = 1;
        for (int i = 1; i <= O; ++i) {
            const int ib = O - i;
            for (int j = 0; j <= i; ++j) {
                const int64_t l = seg - (i - j), r = seg + j + 1;
                R left = 0, right = 0;
                if (j != 0) left = b[ib + j] * (x - knot(l)) / (knot(r - 1) - knot(l));
                if (ib + j != O) right = b[ib + j + 1] * (knot(r) - x) / (knot(r) - knot(l + 1));
                b[ib + j] = left + right;
            }
        }
    }

   private:
    void finish_range(int64_t second_idx) {
        first = knot(order);
        second = knot(second_idx);
    }
};

// Zero-initialised array whose untouched pages cost nothing (calloc): the corner strips of a
// long cyclic axis are almost entirely zero.
template <typename R>
struct ZeroBuf {
    R* p = nullptr;
    size_t count = 0;
    ZeroBuf() = default;
    ZeroBuf(const ZeroBuf&) = delete;
    ZeroBuf& operator=(const ZeroBuf&) = delete;
    ~ZeroBuf() { std::free(p); }
    void assign(size_t n) {
        std::free(p);
        p = static_cast<R*>(std::calloc(n ? n : 1, sizeof(R)));
        if (!p) throw std::bad_alloc();
        count = n;
    }
    R& operator[](size_t i) { return p[i]; }
    const R& operator[](size_t i) const { return p[i]; }
    size_t size() const { return count; }
};

// LU factors of a banded (optionally cyclic) n x n matrix, no pivoting.
// Storage: main band by rows, the cyclic corners as two thin dense strips.
template <typename R>
struct BandFactor {
    int64_t n = 0;
    int p = 0, q = 0;
    bool cyclic = false;
    std::vector<R> band;    // [n][p+q+1]: A(i, j) at band[i*(p+q+1) + j-i+p]
    ZeroBuf<R> right;       // [n][p]:     A(i, n-p+c)  (rows above the band)
    ZeroBuf<R> bottom;      // [q][n]:     A(n-q+r, j)  (columns left of the band)
    int64_t right_rows = 0;   // rows >= right_rows of the right strip were never written
    // Compact form (build_axis_factor): the matrix really has true_n rows; stored row r is true
    // row r for r < head, true row r + (true_n - n) for the last n - head rows, and every true
    // row in between equals stored row head - 1.  true_n == 0: stored as is.
    int64_t true_n = 0, head = 0;
    int64_t full_n() const { return true_n ? true_n : n; }
    int64_t stored_row(int64_t i) const {
        if (!true_n || i < head) return i;
        const int64_t tail0 = true_n - (n - head);
        return i >= tail0 ? i - (true_n - n) : head - 1;
    }
    int64_t bottom_cols = 0;  // columns >= bottom_cols of the bottom strip were never written

    void init(int64_t n_, int p_, int q_, bool cyclic_) {
        n = n_; p = p_; q = q_; cyclic = cyclic_;
        band.assign(static_cast<size_t>(n) * (p + q + 1), R(0));
        if (cyclic) {
            right.assign(static_cast<size_t>(n) * std::max(p, 1));
            bottom.assign(static_cast<size_t>(n) * std::max(q, 1));
            right_rows = bottom_cols = 0;
        }
    }
    bool in_band(int64_t i, int64_t j) const { return j + p >= i && i + q >= j; }
    R& main(int64_t i, int64_t j) { return band[static_cast<size_t>(i) * (p + q + 1) + (j - i + p)]; }
    R main(int64_t i, int64_t j) const { return band[static_cast<size_t>(i) * (p + q + 1) + (j - i + p)]; }
    R& rgt(int64_t i, int64_t j) { return right[static_cast<size_t>(i) * p + (j - (n - p))]; }
    R& bot(int64_t i, int64_t j) { return bottom[static_cast<size_t>(i - (n - q)) * n + j]; }
    // element access valid for band entries and, when cyclic, corner entries
    R& at(int64_t i, int64_t j) {
        if (in_band(i, j)) return main(i, j);
        if (!cyclic) throw std::out_of_range("entry outside the band");
        if (j > i + q) {
            if (j < n - p) throw std::out_of_range("entry outside band and corners");
            right_rows = std::max(right_rows, i + 1);
            return rgt(i, j);
        }
        if (i < n - q) throw std::out_of_range("entry outside band and corners");
        bottom_cols = std::max(bottom_cols, j + 1);
        return bot(i, j);
    }

    // One elimination step (pivot k), right-looking; every entry receives its updates in
    // ascending k, like BandLU.hpp:103-118 / :159-213.
    void step(int64_t k) {
        const R piv = main(k, k);
        const int64_t r_end = std::min<int64_t>(k + p + 1, n);   // band rows   k+1 .. r_end-1
        const int64_t c_end = std::min<int64_t>(k + q + 1, n);   // band cols   k+1 .. c_end-1
        const int64_t rs = cyclic ? std::max<int64_t>(n - q, k + p + 1) : n;  // corner rows rs..n-1
        const int64_t cs = cyclic ? std::max<int64_t>(n - p, k + q + 1) : n;  // corner cols cs..n-1
        for (int64_t i = k + 1; i < r_end; ++i) main(i, k) /= piv;
        if (k < bottom_cols)
            for (int64_t i = rs; i < n; ++i) bot(i, k) /= piv;
        for (int64_t i = k + 1; i < r_end; ++i) {
            const R l = main(i, k);
            for (int64_t j = k + 1; j < c_end; ++j) main(i, j) -= l * main(k, j);
            if (k < right_rows)
                for (int64_t j = cs; j < n; ++j) {
                    const R u = rgt(k, j);
                    if (u != R(0)) at(i, j) -= l * u;
                }
        }
        for (int64_t i = rs; i < n && k < bottom_cols; ++i) {
            const R l = bot(i, k);
            if (l == R(0)) continue;
            for (int64_t j = k + 1; j < c_end; ++j) at(i, j) -= l * main(k, j);
            for (int64_t j = cs; j < n; ++j) main(i, j) -= l * rgt(k, j);
        }
    }

    // Elimination with a fast-forward through the translation-invariant interior of a
    // uniform axis.  Once (a) the rows ahead are bit-identical shifted copies of one another
    // in the ORIGINAL matrix, (b) the corner strips no longer carry anything into the band
    // (their entries in row/column k are exactly zero) and (c) the working state after step k
    // equals the state after step k-1 shifted by one row, every later step in that run
    // reproduces the same bits -- floating point is deterministic -- so its results are
    // copied instead of recomputed.  The factors are bit-identical to the plain loop's
    // (tests/test_abi.py checks that bit for bit); a 2^24-row axis costs O(1000)
    // real steps instead of 1.7e7.
    void factor() {
        const int w = p + q + 1;
        // same[i]: original band row i is a shifted copy of row i-1 and neither touches a corner strip
        std::vector<unsigned char> same(static_cast<size_t>(n), 0);
        if (n > 4 * (p + q + 2)) {
            for (int64_t i = p + 1; i < n - q - 1; ++i) {
                bool eq = true;
                for (int c = 0; c < w && eq; ++c) eq = band[i * w + c] == band[(i - 1) * w + c];
                same[i] = eq;
            }
        }
        std::vector<R> prev_state(static_cast<size_t>(p + 1) * w), cur_state(prev_state.size());
        bool have_prev = false;
        int64_t k = 0;
        while (k + 1 < n) {
            step(k);
            bool can = k + p + 1 < n && same[k + p + 1];
            if (can && cyclic) {
                if (k < right_rows)
                    for (int64_t j = std::max<int64_t>(n - p, k + q + 1); j < n && can; ++j) can = rgt(k, j) == R(0);
                if (k < bottom_cols)
                    for (int64_t i = std::max<int64_t>(n - q, k + p + 1); i < n && can; ++i) can = bot(i, k) == R(0);
            }
            if (!can) { have_prev = false; ++k; continue; }
            for (int r = 0; r <= p; ++r)
                for (int c = 0; c < w; ++c) cur_state[r * w + c] = band[(k + r) * w + c];
            if (have_prev && cur_state == prev_state) {
                // run length: original rows k+p+1 .. stay shift-identical
                int64_t last = k + p + 1;
                while (last + 1 < n && same[last + 1]) ++last;
                const int64_t k_end = last - p;  // steps k+1 .. k_end behave like step k
                if (k_end > k + 1) {
                    for (int64_t r = k + 1; r <= k_end; ++r)
                        for (int c = 0; c < w; ++c) band[r * w + c] = cur_state[c];
                    for (int r = 1; r <= p; ++r)
                        for (int c = 0; c < w; ++c) band[(k_end + r) * w + c] = cur_state[r * w + c];
                    k = k_end + 1;
                    have_prev = false;
                    continue;
                }
            }
            prev_state.swap(cur_state);
            have_prev = true;
            ++k;
        }
    }
};

// Collocation matrix of one axis, factored (build_solver_,
// InterpolationTemplate.hpp:254-446).
//
// Long uniform axes are factored in COMPACT form: away from its two ends the matrix is
// translation invariant (every interior row is the same basis evaluation, :341-360 / :273-280),
// so the elimination settles into a state that repeats bit for bit and the corner strips of a
// periodic axis underflow to exact zeros.  Rows [0, head) and the last n_stored - head rows are
// then factored on a surrogate of n_stored rows assembled from the TRUE axis (true abscissae and
// knots at both ends); every row in between equals row head-1.  The surrogate is accepted only
// if a 128-row window around the cut is bit-identical and the strips died out before it --
// otherwise the full matrix is factored.  tests/test_abi.py compares the expansion with the
// full-size factorisation bit for bit.
constexpr int64_t kCompactRows = 4096;      // surrogate size
constexpr int64_t kCompactMinAxis = 16384;  // axes shorter than this are factored in full

template <typename R>
void assemble_axis_rows(const HostAxis<R>& a, BandFactor<R>& m, int64_t stored, int64_t head, int bw) {
    const int O = a.order;
    const int64_t N = a.n, K = a.K, shift = N - stored;
    m.init(stored, bw, bw, a.periodic);
    R bsv[16] = {0};
    if (a.periodic && a.uniform)  // one evaluation serves every row (:273-280)
        a.basis(O, a.knot(O) + static_cast<R>(1 - O % 2) * a.dx * R(.5), bsv);
    for (int64_t r = 0; r < stored; ++r) {
        const int64_t i = r < head ? r : r + shift;  // true data index of stored row r
        if (!a.periodic && (i == 0 || i == N - 1)) {  // end rows interpolate exactly (:317-329)
            m.main(r, r) = 1;
            continue;
        }
        int64_t seg;
        const bool internal = i > O / 2 && i < N - O / 2 - 1;
        if (a.uniform) {
            seg = a.periodic ? i + O
                             : std::min<int64_t>(K - O - 2, i > O / 2 ? i + (O + 1) / 2 : O);
            if (!a.periodic && (seg <= 2 * O + 1 || seg >= K - 2 * O - 2)) {
                const R x = a.first + static_cast<R>(i) * a.dx;  // (:354-358)
                a.basis(seg, x, bsv);
            }
        } else {
            const R x = a.coords[static_cast<size_t>(i)];
            const int64_t m_coords = static_cast<int64_t>(a.coords.size());
            if (a.periodic) seg = i + O;
            else if (i == 0) seg = O;
            else if (i == m_coords - 1) seg = K - (O + 2);
            else seg = a.span_of(x, i + 1, std::min<int64_t>(K - O - 1, i + O));
            a.basis(seg, x, bsv);
        }
        const int cnt = a.periodic ? (O | 1) : O == 1 ? 1 : (a.uniform && internal) ? (O | 1) : O + 1;
        // placement in stored coordinates (identical to the true ones when nothing is cut out)
        const int64_t row0 = r + (a.periodic ? bw : 0), col0 = seg - (i - r) - O;
        if (row0 < stored && col0 + cnt <= stored && col0 + bw >= row0 && row0 + bw >= col0 + cnt - 1) {
            // no wrap, whole row inside the band: the common case on long axes
            for (int j = 0; j < cnt; ++j) m.main(row0, col0 + j) = bsv[j];
        } else {
            for (int j = 0; j < cnt; ++j) {
                const int64_t row = row0 % stored;
                const int64_t col = (col0 + j) % stored;
                m.at(row, col) = bsv[j];
            }
        }
    }
}

template <typename R>
void build_axis_factor(const HostAxis<R>& a, BandFactor<R>& m, bool allow_compact = true) {
    const int O = a.order;
    const int64_t N = a.n;
    const int bw = a.periodic ? O / 2 : (O == 0 ? 0 : O - 1);
    if (N <= 2 * bw + 1 && a.periodic && bw > 0)
        throw std::invalid_argument("periodic axis too short for this order");
    if (allow_compact && a.uniform && N >= kCompactMinAxis) {
        const int64_t head = kCompactRows / 2, guard = 64;
        assemble_axis_rows(a, m, kCompactRows, head, bw);
        m.factor();
        const int w = m.p + m.q + 1;
        bool steady = !m.cyclic || (m.right_rows <= head - guard && m.bottom_cols <= head - guard);
        for (int64_t r = head - guard + 1; r < head + guard && steady; ++r)
            for (int c = 0; c < w && steady; ++c) steady = m.band[r * w + c] == m.band[(r - 1) * w + c];
        if (steady) {
            m.true_n = N;
            m.head = head;
            return;
        }
    }
    assemble_axis_rows(a, m, N, N, bw);
    m.true_n = 0;
    m.head = 0;
    m.factor();
}

}  // namespace bspl
