// Evaluation kernels (sm_100a): batched InterpolationFunction::operator(),
// derivative() and fused value+gradient.  Replaces the reference's per-point
// BSpline::operator() / derivative_at (BSpline.hpp:305-368, :393-532).
#include "bspl_kernels.h"

namespace bspl {

namespace {

template <typename R, int D>
struct EvalKernelParams {
    AxisParams<R> ax[D];
    const R* coef;
    long long field_stride;
    int n_fields;
    const R* pts;
    R* out;
    long long q;
    int deriv[D];
};

template <int O> constexpr int win() { return 2 * O > 0 ? 2 * O : 1; }

// Per-axis preparation for one query: locate, weights, first control point.
template <typename R, int O, bool GRAD>
__device__ __forceinline__ long long axis_setup(const AxisParams<R>& a, R x, int k, R* w, R* dw) {
    const int span = locate<R, O>(a, x);
    R tk[win<O>()];
    load_knot_window<R, O>(a, span, tk);
    if (GRAD) {
        basis_and_deriv<R, O>(tk, x, w, dw);
    } else {
        if (k == 0) basis_funs<R, O>(tk, x, O, w);
        else deriv_weights<R, O>(tk, x, k, w);
    }
    return static_cast<long long>(span - O) * a.stride;
}

// One query per thread; stencil read straight from the padded global array
// (periodic axes carry O ghost cells, so no index wraps).  Contraction is
// staged: last axis first, so value+gradient costs one gather.
template <typename R, int D, int O, bool GRAD>
__global__ void __launch_bounds__(256) eval_direct_kernel(const EvalKernelParams<R, D> p) {
    constexpr int W = O + 1;
    constexpr int NOUT = GRAD ? D + 1 : 1;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < p.q;
         q += stride) {
        R w[D][W], dw[GRAD ? D : 1][W];
        long long base = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const R x = p.pts[q * D + d];
            base += axis_setup<R, O, GRAD>(p.ax[d], x, GRAD ? 0 : p.deriv[d], w[d],
                                           dw[GRAD ? d : 0]);
        }
        for (int f = 0; f < p.n_fields; ++f) {
            const R* c = p.coef + f * p.field_stride + base;
            R res[NOUT];
            if constexpr (D == 1) {
                R v = R(0), g0 = R(0);
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    const R cv = c[i];
                    v += cv * w[0][i];
                    if (GRAD) g0 += cv * dw[0][i];
                }
                res[0] = v;
                if (GRAD) res[1] = g0;
            } else if constexpr (D == 2) {
                R v = R(0), g0 = R(0), g1 = R(0);
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    const R* row = c + i * p.ax[0].stride;
                    R a = R(0), a1 = R(0);
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        const R cv = row[j];
                        a += cv * w[D - 1][j];
                        if (GRAD) a1 += cv * dw[GRAD ? D - 1 : 0][j];
                    }
                    v += a * w[0][i];
                    if (GRAD) {
                        g0 += a * dw[0][i];
                        g1 += a1 * w[0][i];
                    }
                }
                res[0] = v;
                if (GRAD) { res[1] = g0; res[NOUT - 1] = g1; }
            } else {
                R v = R(0), g0 = R(0), g1 = R(0), g2 = R(0);
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    R bi = R(0), bi1 = R(0), bi2 = R(0);
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        const R* row = c + i * p.ax[0].stride + j * p.ax[D > 1 ? 1 : 0].stride;
                        R a = R(0), a2 = R(0);
#pragma unroll
                        for (int k = 0; k < W; ++k) {
                            const R cv = row[k];
                            a += cv * w[D - 1][k];
                            if (GRAD) a2 += cv * dw[GRAD ? D - 1 : 0][k];
                        }
                        bi += a * w[D > 1 ? 1 : 0][j];
                        if (GRAD) {
                            bi1 += a * dw[GRAD && D > 1 ? 1 : 0][j];
                            bi2 += a2 * w[D > 1 ? 1 : 0][j];
                        }
                    }
                    v += bi * w[0][i];
                    if (GRAD) {
                        g0 += bi * dw[0][i];
                        g1 += bi1 * w[0][i];
                        g2 += bi2 * w[0][i];
                    }
                }
                res[0] = v;
                if (GRAD) { res[1] = g0; res[NOUT > 2 ? 2 : 0] = g1; res[NOUT - 1] = g2; }
            }
            R* o = p.out + (static_cast<long long>(f) * p.q + q) * NOUT;
#pragma unroll
            for (int r = 0; r < NOUT; ++r) o[r] = res[r];
        }
    }
}

// Any dimension up to kMaxDim and any order up to kMaxOrder with run-time loops: the route for 4-D splines and
// for orders 6 and 7, which the reference's templates accept (Interpolation.hpp:17) but the unrolled kernel above is
// not instantiated for.  The stencil is walked with an odometer over the slower axes, the fastest axis contracted
// first; the gradient substitutes the derivative weights one axis at a time.
template <typename R, int O, bool GRAD>
__global__ void __launch_bounds__(128) eval_generic_kernel(const EvalKernelParams<R, kMaxDim> p, int dim) {
    constexpr int W = O + 1;
    const int nout = GRAD ? dim + 1 : 1;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < p.q; q += stride) {
        R w[kMaxDim][W], dw[GRAD ? kMaxDim : 1][W];
        long long base = 0;
        for (int d = 0; d < dim; ++d) {
            const R x = p.pts[q * dim + d];
            base += axis_setup<R, O, GRAD>(p.ax[d], x, GRAD ? 0 : p.deriv[d], w[d], dw[GRAD ? d : 0]);
        }
        int outer = 1;
        for (int d = 0; d + 1 < dim; ++d) outer *= W;
        for (int f = 0; f < p.n_fields; ++f) {
            const R* c = p.coef + f * p.field_stride + base;
            // nested sums, as the unrolled kernels form them: part[s][d] collects, for result s (0: value, 1 + e:
            // derivative along axis e), the terms of the current run of axis d; when axis d + 1 completes a cycle its
            // sum is folded into axis d with that axis' weight.  Rounding error grows with (O + 1) * dim, not (O + 1)^dim.
            R part[GRAD ? kMaxDim + 1 : 1][kMaxDim];
#pragma unroll
            for (int r = 0; r < (GRAD ? kMaxDim + 1 : 1); ++r)
#pragma unroll
                for (int d = 0; d < kMaxDim; ++d) part[r][d] = R(0);
            R v = R(0), g[kMaxDim] = {R(0), R(0), R(0), R(0)};
            for (int t = 0; t < outer; ++t) {
                int idx[kMaxDim] = {0, 0, 0, 0};
                long long off = 0;
                int rem = t;
                for (int d = dim - 2; d >= 0; --d) {
                    idx[d] = rem % W; rem /= W;
                    off += idx[d] * p.ax[d].stride;
                }
                R a = R(0), a1 = R(0);
#pragma unroll
                for (int k = 0; k < W; ++k) {
                    const R cv = c[off + k];
                    a += cv * w[dim - 1][k];
                    if (GRAD) a1 += cv * dw[GRAD ? dim - 1 : 0][k];
                }
                if (dim == 1) {
                    v = a;
                    if (GRAD) g[0] = a1;
                    break;
                }
                // fold the finished innermost sum into axis dim-2, then carry upwards wherever a cycle completed
                const int lastd = dim - 2;
                part[0][lastd] += a * w[lastd][idx[lastd]];
                if (GRAD) {
                    for (int e = 0; e < dim; ++e) {
                        const R src = (e == dim - 1) ? a1 : a;
                        const R wt = (e == lastd) ? dw[GRAD ? lastd : 0][idx[lastd]] : w[lastd][idx[lastd]];
                        part[GRAD ? 1 + e : 0][lastd] += src * wt;
                    }
                }
                for (int d = lastd; d >= 1 && idx[d] == W - 1; --d) {
                    part[0][d - 1] += part[0][d] * w[d - 1][idx[d - 1]];
                    part[0][d] = R(0);
                    if (GRAD) {
                        for (int e = 0; e < dim; ++e) {
                            const R wt = (e == d - 1) ? dw[GRAD ? d - 1 : 0][idx[d - 1]] : w[d - 1][idx[d - 1]];
                            part[GRAD ? 1 + e : 0][d - 1] += part[GRAD ? 1 + e : 0][d] * wt;
                            part[GRAD ? 1 + e : 0][d] = R(0);
                        }
                    }
                }
            }
            if (dim > 1) {
                v = part[0][0];
                if (GRAD)
                    for (int e = 0; e < dim; ++e) g[e] = part[GRAD ? 1 + e : 0][0];
            }
            R* o = p.out + (static_cast<long long>(f) * p.q + q) * nout;
            o[0] = v;
            if (GRAD)
                for (int e = 0; e < dim; ++e) o[1 + e] = g[e];
        }
    }
}

template <typename R, int O>
__global__ void __launch_bounds__(256) locate_generic_kernel(const EvalKernelParams<R, kMaxDim> p, int dim, int32_t* cell) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < p.q; q += stride)
        for (int d = 0; d < dim; ++d) {
            R x = p.pts[q * dim + d];
            cell[q * dim + d] = locate<R, O>(p.ax[d], x) - O;
        }
}

template <typename R, int D, int O>
__global__ void __launch_bounds__(256) locate_kernel(const EvalKernelParams<R, D> p, int32_t* cell) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < p.q;
         q += stride) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            R x = p.pts[q * D + d];
            cell[q * D + d] = locate<R, O>(p.ax[d], x) - O;
        }
    }
}

template <typename R, int D>
EvalKernelParams<R, D> pack(const EvalArgs<R>& a) {
    EvalKernelParams<R, D> p;
    for (int d = 0; d < D; ++d) { p.ax[d] = a.ax[d]; p.deriv[d] = a.deriv[d]; }
    p.coef = a.coef;
    p.field_stride = a.field_stride;
    p.n_fields = a.n_fields;
    p.pts = a.pts;
    p.out = a.out;
    p.q = a.q;
    return p;
}

inline int grid_for(long long q, int block, int max_blocks) {
    long long g = (q + block - 1) / block;
    if (g > max_blocks) g = max_blocks;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

constexpr int kSMs = 148;

template <typename R, int D, int O>
cudaError_t eval_direct_DO(const EvalArgs<R>& a, cudaStream_t s) {
    auto p = pack<R, D>(a);
    const int block = 256;
    const int grid = grid_for(a.q, block, kSMs * 32);
    if (a.mode == kValueGrad) eval_direct_kernel<R, D, O, true><<<grid, block, 0, s>>>(p);
    else eval_direct_kernel<R, D, O, false><<<grid, block, 0, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

template <typename R, int D, int O>
cudaError_t locate_DO(const EvalArgs<R>& a, int32_t* cell, cudaStream_t s) {
    auto p = pack<R, D>(a);
    const int block = 256;
    locate_kernel<R, D, O><<<grid_for(a.q, block, kSMs * 32), block, 0, s>>>(p, cell);
    count_launch();
    return cudaGetLastError();
}

#define BSPL_DISPATCH_ORDER(D_, CALL)                                  \
    switch (a.order) {                                                 \
        case 0: return CALL(D_, 0);                                    \
        case 1: return CALL(D_, 1);                                    \
        case 2: return CALL(D_, 2);                                    \
        case 3: return CALL(D_, 3);                                    \
        case 4: return CALL(D_, 4);                                    \
        case 5: return CALL(D_, 5);                                    \
        default: return cudaErrorInvalidValue;                         \
    }
#define BSPL_DISPATCH(CALL)                                            \
    switch (a.dim) {                                                   \
        case 1: BSPL_DISPATCH_ORDER(1, CALL)                           \
        case 2: BSPL_DISPATCH_ORDER(2, CALL)                           \
        case 3: BSPL_DISPATCH_ORDER(3, CALL)                           \
        default: return cudaErrorInvalidValue;                         \
    }

}  // namespace

template <typename R, int O>
cudaError_t eval_generic_O(const EvalArgs<R>& a, cudaStream_t s) {
    EvalKernelParams<R, kMaxDim> p{};
    for (int d = 0; d < a.dim; ++d) { p.ax[d] = a.ax[d]; p.deriv[d] = a.deriv[d]; }
    p.coef = a.coef; p.field_stride = a.field_stride; p.n_fields = a.n_fields;
    p.pts = a.pts; p.out = a.out; p.q = a.q;
    const int block = 128;
    const int grid = grid_for(a.q, block, kSMs * 32);
    if (a.mode == kValueGrad) eval_generic_kernel<R, O, true><<<grid, block, 0, s>>>(p, a.dim);
    else eval_generic_kernel<R, O, false><<<grid, block, 0, s>>>(p, a.dim);
    count_launch();
    return cudaGetLastError();
}

template <typename R, int O>
cudaError_t locate_generic_O(const EvalArgs<R>& a, int32_t* cell, cudaStream_t s) {
    EvalKernelParams<R, kMaxDim> p{};
    for (int d = 0; d < a.dim; ++d) p.ax[d] = a.ax[d];
    p.pts = a.pts; p.q = a.q;
    locate_generic_kernel<R, O><<<grid_for(a.q, 256, kSMs * 32), 256, 0, s>>>(p, a.dim, cell);
    count_launch();
    return cudaGetLastError();
}

#define BSPL_DISPATCH_GENERIC(CALL)                                    \
    switch (a.order) {                                                 \
        case 0: return CALL(0);                                        \
        case 1: return CALL(1);                                        \
        case 2: return CALL(2);                                        \
        case 3: return CALL(3);                                        \
        case 4: return CALL(4);                                        \
        case 5: return CALL(5);                                        \
        case 6: return CALL(6);                                        \
        case 7: return CALL(7);                                        \
        default: return cudaErrorInvalidValue;                         \
    }

template <typename R>
cudaError_t launch_eval_direct(const EvalArgs<R>& a, cudaStream_t s) {
    if (a.q <= 0) return cudaSuccess;
    if (a.dim < 1 || a.dim > kMaxDim) return cudaErrorInvalidValue;
    if (a.dim > 3 || a.order > 5) {
#define CALL_GEN(O_) eval_generic_O<R, O_>(a, s)
        BSPL_DISPATCH_GENERIC(CALL_GEN)
#undef CALL_GEN
    }
#define CALL_EVAL(D_, O_) eval_direct_DO<R, D_, O_>(a, s)
    BSPL_DISPATCH(CALL_EVAL)
#undef CALL_EVAL
}

template <typename R>
cudaError_t launch_locate(const EvalArgs<R>& a, int32_t* cell, cudaStream_t s) {
    if (a.q <= 0) return cudaSuccess;
    if (a.dim < 1 || a.dim > kMaxDim) return cudaErrorInvalidValue;
    if (a.dim > 3 || a.order > 5) {
#define CALL_GLOC(O_) locate_generic_O<R, O_>(a, cell, s)
        BSPL_DISPATCH_GENERIC(CALL_GLOC)
#undef CALL_GLOC
    }
#define CALL_LOC(D_, O_) locate_DO<R, D_, O_>(a, cell, s)
    BSPL_DISPATCH(CALL_LOC)
#undef CALL_LOC
}

template cudaError_t launch_eval_direct<double>(const EvalArgs<double>&, cudaStream_t);
template cudaError_t launch_eval_direct<float>(const EvalArgs<float>&, cudaStream_t);
template cudaError_t launch_locate<double>(const EvalArgs<double>&, int32_t*, cudaStream_t);
template cudaError_t launch_locate<float>(const EvalArgs<float>&, int32_t*, cudaStream_t);

}  // namespace bspl
