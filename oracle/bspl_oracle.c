/* TEST INFRASTRUCTURE ONLY -- see bspl_oracle.h.  Plain-C restatement of the
 * reference's evaluate + control-point-solve path.  Compiled with
 * -ffp-contract=off so every a*b+c rounds twice, like the reference built for
 * baseline x86-64 (no FMA).  Citations: <file>:<lines> in the reference. */
#include "bspl_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXW (BSPLO_MAXO + 1)

static int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
static int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }

/* ---- knots ------------------------------------------------------------- */

/* BSpline::load_knots, BSpline.hpp:217-227 */
static void set_range(bsplo_axis* a) {
    int O = a->order;
    a->first = a->t[O];
    a->second = a->t[a->K - O - (2 - O % 2)];
}

/* create_knot_vector_ (uniform overload), Interpolation.hpp:322-362,
 * with INTP_PERIODIC_NO_DUMMY_POINT (:331-333). */
static void knots_uniform(bsplo_axis* a, double lo, double hi) {
    int O = a->order;
    int64_t n = a->n + (a->periodic ? 1 : 0);
    a->dx = (hi - lo) / (double)(n - 1);
    int64_t extra = a->periodic ? 2 * O + (1 - O % 2) : O + 1;
    a->K = n + extra;
    a->t = (double*)malloc(sizeof(double) * (size_t)a->K);
    for (int64_t i = 0; i < a->K; ++i) a->t[i] = lo;
    if (a->periodic) {
        for (int64_t i = 0; i < a->K; ++i)
            a->t[i] = lo + ((double)i - .5 * (double)extra) * a->dx;
    } else {
        for (int64_t i = O + 1; i < a->K - O - 1; ++i)
            a->t[i] = lo + ((double)i - .5 * (double)extra) * a->dx;
        for (int64_t i = a->K - O - 1; i < a->K; ++i) a->t[i] = hi;
    }
    a->uniform = 1;
    a->coords = NULL;
    set_range(a);
}

/* create_knot_vector_ (iterator-pair overload), Interpolation.hpp:365-464 */
static void knots_nonuniform(bsplo_axis* a, const double* c) {
    int O = a->order;
    int64_t n = a->n + (a->periodic ? 1 : 0); /* number of abscissae given */
    a->K = a->periodic ? n + 2 * O + (1 - O % 2) : n + O + 1;
    a->t = (double*)calloc((size_t)a->K, sizeof(double));
    a->coords = (double*)malloc(sizeof(double) * (size_t)n);
    double* xs = a->t;
    double* ic = a->coords;
    int64_t m = 0; /* ic fill count */
    if (a->periodic) {
        const double* it = c;
        ic[m++] = *it;
        for (int64_t i = O + 1; i < O + n; ++i) {
            double present = *(++it);
            xs[i] = (O % 2 == 0) ? .5 * (ic[m - 1] + present) : present;
            ic[m++] = present;
        }
        double period = ic[m - 1] - ic[0];
        for (int64_t i = 0; i < O + 1; ++i) {
            xs[i] = xs[n + i - 1] - period;
            xs[a->K - i - 1] = xs[a->K - i - n] + period;
        }
    } else {
        const double* it = c;
        double l_knot = *it;
        for (int64_t i = 0; i < O + 1; ++i) xs[i] = l_knot;
        ic[m++] = l_knot;
        double window_sum = 0;
        for (int64_t i = 1; i < O; ++i) {
            ic[m++] = *(++it);
            window_sum += ic[i];
        }
        for (int64_t i = O + 1; i < n; ++i) {
            ic[m++] = *(++it);
            window_sum += ic[i - 1];
            xs[i] = window_sum / (double)O;
            window_sum -= ic[i - O];
        }
        double r_knot = *(++it);
        for (int64_t i = n; i < n + O + 1; ++i) xs[i] = r_knot;
        ic[m++] = r_knot;
    }
    a->uniform = 0;
    a->dx = 0;
    set_range(a);
}

/* ---- locate + basis ---------------------------------------------------- */

/* get_knot_iter, BSpline.hpp:125-157: wraps x (periodic), accepts the hint if
 * t[h] <= x < t[h+1], else --upper_bound(t+O+1, t+last+1, x). */
static int64_t knot_index(const bsplo_axis* a, double* x, int64_t hint, int64_t last) {
    const double* t = a->t;
    if (a->periodic) {
        double period = a->second - a->first;
        *x = a->first + fmod(*x - a->first, period) + (*x < a->first ? period : 0.);
    }
    if (t[hint] <= *x && t[hint + 1] > *x) return hint;
    int64_t lo = a->order + 1, hi = last + 1; /* first idx in [lo,hi) with t > x */
    while (lo < hi) {
        int64_t mid = lo + (hi - lo) / 2;
        if (!(*x < t[mid])) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

/* add_hint_for_spline, Interpolation.hpp:277-295 */
static int64_t span_hint(const bsplo_axis* a, double x) {
    int O = a->order;
    if (!a->uniform) return O;
    double v = (x - a->first) / a->dx - (a->periodic ? 1. : .5 * (double)(O + 1));
    v = ceil(v > 0. ? v : 0.);
    int64_t h = (v >= 9.0e18) ? INT64_MAX / 2 : (int64_t)v;
    return imin(a->K - O - 2, h + O);
}

static int64_t locate(const bsplo_axis* a, double* x) {
    return knot_index(a, x, span_hint(a, *x), a->K - a->order - 2);
}

/* base_spline_value, BSpline.hpp:83-111.  seg = index of the segment's left
 * knot; lower spline_order leaves values right-aligned. */
static void basis(const bsplo_axis* a, int64_t seg, double x, int spline_order, double* b) {
    int O = a->order;
    const double* t = a->t;
    for (int i = 0; i <= O; ++i) b[i] = 0.;
    b[O] = 1.;
    for (int i = 1; i <= spline_order; ++i) {
        int ib = O - i;
        for (int j = 0; j <= i; ++j) {
            int64_t l = seg - (i - j), r = seg + j + 1;
            double left = (j == 0) ? 0. : b[ib + j] * (x - t[l]) / (t[r - 1] - t[l]);
            double right =
                (ib + j == O) ? 0. : b[ib + j + 1] * (t[r] - x) / (t[r] - t[l + 1]);
            b[ib + j] = left + right;
        }
    }
}

/* ---- evaluation -------------------------------------------------------- */

static int64_t ipow(int64_t b, int e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }

/* control point at multi-index c0+digits, periodic axes taken modulo n
 * (BSpline.hpp:357-360; the cell layout bakes the same wrap in, :721-728). */
static double ctrl_at(const bsplo_spline* s, const int64_t* c0, const int* dig) {
    int64_t lin = 0;
    for (int d = 0; d < s->dim; ++d) {
        int64_t i = c0[d] + dig[d];
        if (s->ax[d].periodic) i %= s->ax[d].n;
        lin = lin * s->ax[d].n + i;
    }
    return s->ctrl[lin];
}

/* BSpline::operator(), BSpline.hpp:305-335 (INTP_CELL_LAYOUT summation order:
 * flat stencil index with axis 0 as the fastest digit; coef = C, then *= b_d). */
static double eval_one(const bsplo_spline* s, const double* pt) {
    int D = s->dim, O = s->order;
    double x[BSPLO_MAXD], b[BSPLO_MAXD][MAXW];
    int64_t c0[BSPLO_MAXD];
    for (int d = 0; d < D; ++d) {
        x[d] = pt[d];
        int64_t seg = locate(&s->ax[d], &x[d]);
        basis(&s->ax[d], seg, x[d], O, b[d]);
        c0[d] = seg - O;
    }
    int64_t total = ipow(O + 1, D);
    double v = 0.;
    for (int64_t i = 0; i < total; ++i) {
        int dig[BSPLO_MAXD];
        int64_t ci = i;
        for (int d = 0; d < D; ++d) { dig[d] = (int)(ci % (O + 1)); ci /= (O + 1); }
        double coef = ctrl_at(s, c0, dig);
        for (int d = 0; d < D; ++d) coef *= b[d][dig[d]];
        v += coef;
    }
    return v;
}

/* BSpline::derivative_at, BSpline.hpp:393-532 */
static double deriv_one(const bsplo_spline* s, const double* pt, const int* dv) {
    int D = s->dim, O = s->order;
    int so[BSPLO_MAXD];
    for (int d = 0; d < D; ++d) {
        if (dv[d] > O) return 0.; /* :404-407 */
        so[d] = O - dv[d];
    }
    double x[BSPLO_MAXD], b[BSPLO_MAXD][MAXW];
    int64_t seg[BSPLO_MAXD], c0[BSPLO_MAXD];
    for (int d = 0; d < D; ++d) {
        x[d] = pt[d];
        seg[d] = locate(&s->ax[d], &x[d]);
        basis(&s->ax[d], seg[d], x[d], so[d], b[d]);
        c0[d] = seg[d] - O;
    }
    int64_t total = ipow(O + 1, D);
    double* lc = (double*)malloc(sizeof(double) * (size_t)total * 2);
    double* lw = lc + total;
    /* local arrays are row-major Mesh<dim>(O+1): stride of axis d = (O+1)^(D-1-d) */
    int64_t stride[BSPLO_MAXD];
    for (int d = 0; d < D; ++d) stride[d] = ipow(O + 1, D - 1 - d);
    for (int64_t i = 0; i < total; ++i) { /* :441-455 */
        int dig[BSPLO_MAXD];
        int64_t ci = i, li = 0;
        double coef = 1.;
        for (int d = 0; d < D; ++d) { dig[d] = (int)(ci % (O + 1)); ci /= (O + 1); }
        for (int d = 0; d < D; ++d) { coef *= b[d][dig[d]]; li += dig[d] * stride[d]; }
        lw[li] = coef;
        lc[li] = ctrl_at(s, c0, dig);
    }
    for (int d = 0; d < D; ++d) { /* :489-520 */
        if (so[d] == O) continue;
        const double* t = s->ax[d].t + seg[d];
        int64_t hs = total / (O + 1);
        for (int64_t i = 0; i < hs; ++i) {
            int64_t ci = i, base = 0;
            for (int dd = 0; dd < D; ++dd) {
                if (dd == d) continue;
                base += (ci % (O + 1)) * stride[dd];
                ci /= (O + 1);
            }
            double* it = lc + base;
            int64_t st = stride[d];
            for (int k = O; k > so[d]; --k)
                for (int j = k; j > 0; --j)
                    it[(O + j - k) * st] = (double)k *
                                           (it[(O + j - k) * st] - it[(O + j - k - 1) * st]) /
                                           (t[j] - t[j - k]);
        }
    }
    double v = 0.;
    for (int64_t i = 0; i < total; ++i) v += lw[i] * lc[i]; /* :524-529 */
    free(lc);
    return v;
}

void bsplo_spans(const bsplo_spline* s, const double* pts, int64_t q, int64_t* out) {
    for (int64_t i = 0; i < q; ++i)
        for (int d = 0; d < s->dim; ++d) {
            double x = pts[i * s->dim + d];
            out[i * s->dim + d] = locate(&s->ax[d], &x) - s->order;
        }
}

void bsplo_eval(const bsplo_spline* s, const double* pts, int64_t q, double* out, int nthreads) {
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
    for (int64_t i = 0; i < q; ++i) out[i] = eval_one(s, pts + i * s->dim);
}

void bsplo_deriv(const bsplo_spline* s, const double* pts, int64_t q, const int* deriv,
                 double* out, int nthreads) {
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
    for (int64_t i = 0; i < q; ++i) out[i] = deriv_one(s, pts + i * s->dim, deriv);
}

/* ---- band LU ----------------------------------------------------------- */

/* storage accessors: BandMatrix.hpp:50-60, :122-146 */
#define MAINB(a, i, j) ((a)->band[(j) * (1 + (a)->p + (a)->q) + ((i) + (a)->q - (j))])
static double* side(bsplo_axis* a, int64_t i, int64_t j) {
    return (j > i + a->q) ? &a->right[i * a->p + (j + a->p - a->n)]
                          : &a->bottom[j * a->q + (i + a->q - a->n)];
}
static double* elem(bsplo_axis* a, int64_t i, int64_t j) {
    if (a->periodic && (j > i + a->q || i > j + a->p)) return side(a, i, j);
    return &MAINB(a, i, j);
}

static void alloc_matrix(bsplo_axis* a, int64_t bw) {
    a->p = a->q = bw;
    a->band = (double*)calloc((size_t)(a->n * (1 + 2 * bw)), sizeof(double));
    a->right = a->bottom = NULL;
    if (a->periodic) {
        a->right = (double*)calloc((size_t)imax(1, (a->n - bw - 1) * bw), sizeof(double));
        a->bottom = (double*)calloc((size_t)imax(1, (a->n - bw - 1) * bw), sizeof(double));
    }
}

/* BandLU<BandMatrix>::compute_impl, BandLU.hpp:103-118 */
static void lu_band(bsplo_axis* a) {
    int64_t n = a->n, p = a->p, q = a->q;
    for (int64_t k = 0; k < n - 1; ++k) {
        for (int64_t i = k + 1; i < imin(k + p + 1, n); ++i) MAINB(a, i, k) /= MAINB(a, k, k);
        for (int64_t j = k + 1; j < imin(k + q + 1, n); ++j)
            for (int64_t i = k + 1; i < imin(k + p + 1, n); ++i)
                MAINB(a, i, j) -= MAINB(a, i, k) * MAINB(a, k, j);
    }
}

/* BandLU<ExtendedBandMatrix>::compute_impl, BandLU.hpp:159-213 */
static void lu_cyclic(bsplo_axis* a) {
    int64_t n = a->n, p = a->p, q = a->q;
    for (int64_t k = 0; k < n - 1; ++k) {
        for (int64_t i = k + 1; i < imin(k + p + 1, n); ++i) MAINB(a, i, k) /= MAINB(a, k, k);
        for (int64_t i = imax(n - q, k + p + 1); i < n; ++i) *side(a, i, k) /= MAINB(a, k, k);
        for (int64_t j = k + 1; j < imin(k + q + 1, n); ++j)
            for (int64_t i = k + 1; i < imin(k + p + 1, n); ++i)
                MAINB(a, i, j) -= MAINB(a, i, k) * MAINB(a, k, j);
        for (int64_t i = k + 1; i < imin(k + p + 1, n); ++i)
            for (int64_t j = imax(n - p, k + q + 1); j < n; ++j)
                *elem(a, i, j) -= MAINB(a, i, k) * *side(a, k, j);
        for (int64_t j = k + 1; j < imin(k + q + 1, n); ++j)
            for (int64_t i = imax(n - q, k + p + 1); i < n; ++i)
                *elem(a, i, j) -= *side(a, i, k) * MAINB(a, k, j);
        if (k < imax(n - p - 1, n - q - 1))
            for (int64_t i = imax(n - q, k + p + 1); i < n; ++i)
                for (int64_t j = imax(n - p, k + q + 1); j < n; ++j)
                    MAINB(a, i, j) -= *side(a, i, k) * *side(a, k, j);
    }
}

/* solve_in_place_impl on a strided line: BandLU.hpp:120-143 (band) and
 * :215-259 (cyclic) */
static void solve_line(const bsplo_axis* ac, double* x, int64_t st) {
    bsplo_axis* a = (bsplo_axis*)ac;
    int64_t n = a->n, p = a->p, q = a->q;
    if (!a->periodic) {
        for (int64_t j = 0; j < n; ++j)
            for (int64_t i = j + 1; i < imin(j + p + 1, n); ++i)
                x[i * st] -= MAINB(a, i, j) * x[j * st];
        for (int64_t j = n - 1; j >= 0; --j) {
            x[j * st] /= MAINB(a, j, j);
            for (int64_t i = (j < q ? 0 : j - q); i < j; ++i)
                x[i * st] -= MAINB(a, i, j) * x[j * st];
        }
    } else {
        for (int64_t j = 0; j < n; ++j) {
            for (int64_t i = j + 1; i < imin(j + p + 1, n); ++i)
                x[i * st] -= MAINB(a, i, j) * x[j * st];
            if (j < n - p - 1)
                for (int64_t i = imax(n - q, j + p + 1); i < n; ++i)
                    x[i * st] -= *side(a, i, j) * x[j * st];
        }
        for (int64_t j = n - 1; j >= 0; --j) {
            x[j * st] /= MAINB(a, j, j);
            for (int64_t i = (j < q ? 0 : j - q); i < j; ++i)
                x[i * st] -= MAINB(a, i, j) * x[j * st];
            if (j > n - p - 1)
                for (int64_t i = 0; i < j - q; ++i) x[i * st] -= *side(a, i, j) * x[j * st];
        }
    }
}

/* build_solver_, InterpolationTemplate.hpp:254-446 */
static void build_solver(bsplo_axis* a) {
    int O = a->order;
    int64_t N = a->n, K = a->K;
    int64_t bw = a->periodic ? O / 2 : (O == 0 ? 0 : O - 1);
    alloc_matrix(a, bw);
    double bsv[MAXW];
    memset(bsv, 0, sizeof bsv);
    if (a->periodic && a->uniform) /* :273-280 */
        basis(a, O, a->t[O] + (double)(1 - O % 2) * a->dx * .5, O, bsv);
    for (int64_t i = 0; i < N; ++i) {
        if (!a->periodic && (i == 0 || i == N - 1)) { /* :317-329 */
            MAINB(a, i, i) = 1.;
            continue;
        }
        int64_t knot_ind;
        int is_internal = i > O / 2 && i < N - O / 2 - 1;
        if (a->uniform) { /* :341-360 */
            knot_ind = a->periodic ? i + O
                                   : imin(K - O - 2, i > O / 2 ? i + (O + 1) / 2 : O);
            if (!a->periodic && (knot_ind <= 2 * O + 1 || knot_ind >= K - 2 * O - 2)) {
                double x = a->first + (double)i * a->dx;
                basis(a, knot_ind, x, O, bsv);
            }
        } else { /* :361-380 */
            double x = a->coords[i];
            int64_t ncoord = a->n + (a->periodic ? 1 : 0);
            if (a->periodic) knot_ind = i + O;
            else if (i == 0) knot_ind = O;
            else if (i == ncoord - 1) knot_ind = K - (O + 2);
            else knot_ind = knot_index(a, &x, i + 1, imin(K - O - 1, i + O));
            basis(a, knot_ind, x, O, bsv);
        }
        int64_t s_num = a->periodic ? (O | 1)
                        : O == 1    ? 1
                        : (a->uniform && is_internal) ? (O | 1) : O + 1; /* :383-386 */
        for (int64_t j = 0; j < s_num; ++j) { /* :387-398 */
            int64_t row = (i + (a->periodic ? bw : 0)) % N;
            int64_t col = (knot_ind - O + j) % N;
            *elem(a, row, col) = bsv[j];
        }
    }
    if (a->periodic) lu_cyclic(a); else lu_band(a);
}

/* ---- control-point solve ---------------------------------------------- */

/* solve_for_control_points_, InterpolationTemplate.hpp:448-580.  The
 * reference solves along the last axis and rotates the array after every
 * sweep (:493-545); solving the same lines in place along a strided axis
 * performs the identical per-line arithmetic, in the same axis order
 * (solvers_[D-1], then D-2, ..., 0; :515). */
int bsplo_interpolate(bsplo_spline* s, const double* f, int nthreads) {
    int D = s->dim, O = s->order;
    for (int d = 0; d < D; ++d) if (!s->ax[d].band) return 1;
    int64_t total = s->size;
    double* w = s->ctrl;
    for (int64_t lin = 0; lin < total; ++lin) { /* :451-462 */
        int64_t rem = lin, dst = 0, mul = 1;
        for (int d = D - 1; d >= 0; --d) {
            int64_t n = s->ax[d].n, i = rem % n;
            rem /= n;
            if (s->ax[d].periodic) i = (i + n + O / 2) % n;
            dst += i * mul;
            mul *= n;
        }
        w[dst] = f[lin];
    }
    for (int d = D - 1; d >= 0; --d) {
        int64_t n = s->ax[d].n, st = 1;
        for (int e = d + 1; e < D; ++e) st *= s->ax[e].n;
        int64_t lines = total / n;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) schedule(static)
        for (int64_t l = 0; l < lines; ++l) {
            int64_t outer = l / st, inner = l % st;
            solve_line(&s->ax[d], w + outer * n * st + inner, st);
        }
    }
    return 0;
}

/* ---- construction ------------------------------------------------------ */

static void finish(bsplo_spline* s) {
    s->size = 1;
    for (int d = 0; d < s->dim; ++d) s->size *= s->ax[d].n;
    s->ctrl = (double*)calloc((size_t)s->size, sizeof(double));
}

bsplo_spline* bsplo_create(int dim, int order, const int64_t* n, const int* periodic,
                           const double* lo, const double* hi, const double* const* coords,
                           int with_solver) {
    if (dim < 1 || dim > BSPLO_MAXD || order < 0 || order > BSPLO_MAXO) return NULL;
    for (int d = 0; d < dim; ++d)
        if (n[d] < order + 1 || n[d] < 2) return NULL; /* the reference has no such guard */
    bsplo_spline* s = (bsplo_spline*)calloc(1, sizeof *s);
    s->dim = dim;
    s->order = order;
    for (int d = 0; d < dim; ++d) {
        bsplo_axis* a = &s->ax[d];
        a->order = order;
        a->periodic = periodic[d] != 0;
        a->n = n[d];
        if (coords && coords[d]) knots_nonuniform(a, coords[d]);
        else knots_uniform(a, lo[d], hi[d]);
        if (with_solver) build_solver(a);
    }
    finish(s);
    return s;
}

bsplo_spline* bsplo_from_knots(int dim, int order, const int64_t* n_ctrl, const int* periodic,
                               const double* const* knots, const int64_t* n_knots,
                               const double* ctrl) {
    if (dim < 1 || dim > BSPLO_MAXD || order < 0 || order > BSPLO_MAXO) return NULL;
    bsplo_spline* s = (bsplo_spline*)calloc(1, sizeof *s);
    s->dim = dim;
    s->order = order;
    for (int d = 0; d < dim; ++d) {
        bsplo_axis* a = &s->ax[d];
        a->order = order;
        a->periodic = periodic[d] != 0;
        a->n = n_ctrl[d];
        a->K = n_knots[d];
        a->t = (double*)malloc(sizeof(double) * (size_t)a->K);
        memcpy(a->t, knots[d], sizeof(double) * (size_t)a->K);
        a->uniform = 0; /* hint = order, BSpline.hpp:376-382 */
        /* range from this ctor: (t[O], t[K-O-1]), BSpline.hpp:200-202 */
        a->first = a->t[order];
        a->second = a->t[a->K - order - 1];
    }
    finish(s);
    memcpy(s->ctrl, ctrl, sizeof(double) * (size_t)s->size);
    return s;
}

void bsplo_destroy(bsplo_spline* s) {
    if (!s) return;
    for (int d = 0; d < s->dim; ++d) {
        free(s->ax[d].t);
        free(s->ax[d].coords);
        free(s->ax[d].band);
        free(s->ax[d].right);
        free(s->ax[d].bottom);
    }
    free(s->ctrl);
    free(s);
}

int bsplo_band_solve(int64_t n, int64_t p, int64_t q, int cyclic, const double* A, double* x) {
    bsplo_axis a;
    memset(&a, 0, sizeof a);
    a.n = n;
    a.periodic = cyclic != 0;
    a.p = p;
    a.q = q;
    a.band = (double*)calloc((size_t)(n * (1 + p + q)), sizeof(double));
    if (cyclic) {
        a.right = (double*)calloc((size_t)imax(1, (n - q - 1) * p), sizeof(double));
        a.bottom = (double*)calloc((size_t)imax(1, (n - p - 1) * q), sizeof(double));
    }
    for (int64_t i = 0; i < n; ++i)
        for (int64_t j = 0; j < n; ++j) {
            int in_main = (j + p >= i) && (i + q >= j);
            int in_right = cyclic && j >= imax(n - p, i + q + 1);
            int in_bottom = cyclic && i >= imax(n - q, j + p + 1);
            if (in_main || in_right || in_bottom) *elem(&a, i, j) = A[i * n + j];
        }
    if (cyclic) lu_cyclic(&a); else lu_band(&a);
    solve_line(&a, x, 1);
    free(a.band);
    free(a.right);
    free(a.bottom);
    return 0;
}
