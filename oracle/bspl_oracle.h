/* TEST INFRASTRUCTURE ONLY -- a plain-C, CPU restatement of the reference's hot
 * path (12ff54e/BSplineInterpolation, header-only C++).  Imported only by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; the product
 * (bsplineinterpolation_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 * (1) every Mathematica golden vector the reference's own tests hold for this
 * path (tests/golden/reference_vectors.json, extracted from
 * test/src/interpolation-test.cpp and test/src/bspline-test.cpp) and (2)
 * outputs of the unmodified reference headers compiled here into oracle/_ref
 * (tests/golden/ref_outputs.npz + live comparison when oracle/_ref exists).
 *
 * All file:line citations are relative to the reference checkout.
 */
#ifndef BSPL_ORACLE_H
#define BSPL_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSPLO_MAXD 4
#define BSPLO_MAXO 7

typedef struct {
    int order;
    int periodic;
    int uniform;
    int64_t n;      /* data points == control points on this axis              */
    int64_t K;      /* number of knots                                          */
    double* t;      /* knots                                                    */
    double first;   /* range().first  = t[O]                  BSpline.hpp:224   */
    double second;  /* range().second = t[K-O-(2-O%2)]        BSpline.hpp:225   */
    double dx;      /* uniform spacing                   Interpolation.hpp:335  */
    double* coords; /* data abscissae (non-uniform only), n (+1 periodic)       */
    /* band LU of the collocation matrix (InterpolationTemplate.hpp:254-446)    */
    int64_t p, q;
    double* band;   /* n x (1+p+q): A(i,j) = band[j*(1+p+q) + i+q-j]            */
    double* right;  /* (n-q-1) x p : A(i,j) = right[i*p + j+p-n]   (periodic)   */
    double* bottom; /* (n-p-1) x q : A(i,j) = bottom[j*q + i+q-n]  (periodic)   */
} bsplo_axis;

typedef struct {
    int dim, order;
    bsplo_axis ax[BSPLO_MAXD];
    double* ctrl; /* plain control points, row-major, same shape as the mesh */
    int64_t size;
} bsplo_spline;

/* Build knots (+ collocation LU when with_solver != 0).  coords[d] == NULL ->
 * uniform axis on [lo[d], hi[d]]; else n[d] (+1 if periodic) abscissae. */
bsplo_spline* bsplo_create(int dim, int order, const int64_t* n, const int* periodic,
                           const double* lo, const double* hi, const double* const* coords,
                           int with_solver);
/* Spline straight from knots + control points (the BSpline ctor,
 * BSpline.hpp:188-210): n_ctrl[d] control points, n_knots[d] knots per axis. */
bsplo_spline* bsplo_from_knots(int dim, int order, const int64_t* n_ctrl, const int* periodic,
                               const double* const* knots, const int64_t* n_knots,
                               const double* ctrl);
void bsplo_destroy(bsplo_spline* s);

/* Control-point solve (InterpolationTemplate.hpp:448-580): f is the row-major
 * mesh; the result is stored in s->ctrl.  nthreads > 1 splits lines (OpenMP). */
int bsplo_interpolate(bsplo_spline* s, const double* f, int nthreads);

/* span - order per axis for each query; pts is [q][dim]. */
void bsplo_spans(const bsplo_spline* s, const double* pts, int64_t q, int64_t* out);
void bsplo_eval(const bsplo_spline* s, const double* pts, int64_t q, double* out, int nthreads);
void bsplo_deriv(const bsplo_spline* s, const double* pts, int64_t q, const int* deriv,
                 double* out, int nthreads);

/* Raw band solvers (BandLU.hpp): dense row-major n x n input, rhs solved in
 * place.  cyclic != 0 selects the bordered ("extended") variant. */
int bsplo_band_solve(int64_t n, int64_t p, int64_t q, int cyclic, const double* a, double* x);

#ifdef __cplusplus
}
#endif
#endif
