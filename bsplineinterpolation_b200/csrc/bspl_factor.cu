// Collocation matrix and band LU of a long NON-UNIFORM axis, built on the device:
//   build_solver_              InterpolationTemplate.hpp:254-446   (rows: basis values at the data abscissae)
//   BandLU::compute_impl       BandLU.hpp:103-118                  (Doolittle in band, no pivoting)
// bspl_host.h does the same on the host, serially: fine for the short axes of 2-D / 3-D meshes and for long
// UNIFORM axes, whose translation-invariant interior is factored in O(1) (compact factors).  A long non-uniform
// axis has no such shortcut -- every row is its own basis evaluation and its own elimination step -- and the
// serial host code costs about a second per 2^24 rows, plus the upload of the factor tables.
//
//   assemble_rows_kernel   one thread per data abscissa: span (the reference's hint + upper_bound), Cox-de Boor
//                          triangle on the explicit knots, the O+1 values dropped into the band row.  Same
//                          operations, same order, same rounding as HostAxis::basis (every product, quotient
//                          and sum rounded separately), so the matrix is bit-identical to the host's.
//   chunk_lu_kernel        the elimination itself is a recurrence along the rows (row k needs rows k-p .. k-1
//                          final), but a contractive one: the influence of the state decays geometrically
//                          (like the substitution sweeps, bspl_solve.cu).  The rows are cut into chunks; the
//                          thread of a chunk starts `window` rows early from the untouched matrix and, by the
//                          time it reaches its own rows, reproduces the sequential state.  A right-looking step
//                          on a (p+1)-row window held in registers, updates in ascending pivot order, multiply
//                          and subtract rounded separately, IEEE division -- BandFactor::step's arithmetic.
//   Every chunk also runs kCheckRows rows into its successor's range; check_overlap_kernel compares those rows
//   with the successor's own, bit for bit.  A single differing bit refuses the result and the caller factors
//   on the host instead: the device path is taken only where it provably reproduces the sequential factors
//   at the seams, and tests/test_gpu_parity.py compares whole solves with the reference's sequential LU.
//
// Periodic (cyclic) axes: the band part of the bordered LU (BandLU.hpp:159-213) does not depend on the border at all,
// and the border -- the two corner strips, which die out geometrically away from the corner, and what they subtract
// from the P x P corner block -- depends on the first few hundred pivots only.  device_factor (bspl_capi.cu) runs those
// pivots on the host, on a 4 096-row surrogate assembled from the true axis (BandFactor, bspl_host.h), hands the
// corrected corner block to the band assembled here, and takes strips and head rows from the host, everything else
// from the chunked elimination below.
#include "bspl_kernels.h"

namespace bspl {

namespace {

template <typename R> __device__ __forceinline__ R div_rn_(R a, R b);
template <> __device__ __forceinline__ double div_rn_<double>(double a, double b) { return __ddiv_rn(a, b); }
template <> __device__ __forceinline__ float div_rn_<float>(float a, float b) { return __fdiv_rn(a, b); }

constexpr int kFactorMaxOrder = 5;
constexpr int kCheckRows = 8;

// HostAxis::span_of
template <typename R>
__device__ __forceinline__ long long span_of(const R* __restrict__ t, int order, R x, long long hint, long long last) {
    if (t[hint] <= x && t[hint + 1] > x) return hint;
    long long a = order + 1, b = last + 1;
    while (a < b) {
        const long long mid = a + (b - a) / 2;
        if (!(x < t[mid])) a = mid + 1; else b = mid;
    }
    return a - 1;
}

// HostAxis::basis (base_spline_value, BSpline.hpp:83-111)
template <typename R>
__device__ __forceinline__ void basis_at(const R* __restrict__ t, int O, long long seg, R x, R* b) {
    using A = Arith<R>;
    for (int i = 0; i <= O; ++i) b[i] = R(0);
    b[O] = R(1);
    for (int i = 1; i <= O; ++i) {
        const int ib = O - i;
        for (int j = 0; j <= i; ++j) {
            const long long l = seg - (i - j), r = seg + j + 1;
            R left = R(0), right = R(0);
            if (j != 0) left = div_rn_<R>(A::mul(b[ib + j], A::sub(x, t[l])), A::sub(t[r - 1], t[l]));
            if (ib + j != O) right = div_rn_<R>(A::mul(b[ib + j + 1], A::sub(t[r], x)), A::sub(t[r], t[l + 1]));
            b[ib + j] = A::add(left, right);
        }
    }
}

// assemble_axis_rows, non-uniform branches: band[i][j - i + bw] = A(i, j).  The band is zeroed by the caller.
// Periodic axes (bordered form, InterpolationTemplate.hpp:383-392): abscissa i fills matrix row i + bw at columns
// i .. i + cnt - 1; rows and columns that wrap around belong to the corner strips, which the caller builds on the
// host (device_factor in bspl_capi.cu), and are skipped here.
template <typename R>
__global__ void __launch_bounds__(256) assemble_rows_kernel(const R* __restrict__ coords, const R* __restrict__ t, int O,
                                                            long long n, long long K, int bw, int periodic,
                                                            R* __restrict__ band) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int w = 2 * bw + 1;
    R bsv[kFactorMaxOrder + 1];
    if (periodic) {
        const long long row0 = i + bw;
        if (row0 >= n) return;
        basis_at<R>(t, O, i + O, coords[i], bsv);
        const int cnt = O | 1;
        for (int j = 0; j < cnt; ++j) {
            const long long col = i + j, c = col - row0 + bw;
            if (col < n && c >= 0 && c < w) band[row0 * w + c] = bsv[j];
        }
        return;
    }
    R* row = band + i * w;
    if (i == 0 || i == n - 1) {  // end rows interpolate exactly (:317-329)
        row[bw] = R(1);
        return;
    }
    const R x = coords[i];
    const long long last = (K - O - 1 < i + O) ? K - O - 1 : i + O;
    const long long seg = span_of<R>(t, O, x, i + 1, last);
    basis_at<R>(t, O, seg, x, bsv);
    const int cnt = O == 1 ? 1 : O + 1;
    const long long col0 = seg - O;
    for (int j = 0; j < cnt; ++j) {
        const long long c = col0 + j - i + bw;
        if (c >= 0 && c < w) row[c] = bsv[j];
    }
}

// Right-looking band LU of rows [a, b) of chunk c, warmed up from max(0, a - window).  The window of p+1 rows
// lives in registers: win[r][.] is row k + r of the working matrix while pivot k is processed.
template <typename R, int P>
__global__ void __launch_bounds__(128) chunk_lu_kernel(const R* __restrict__ band, long long n, long long first_row,
                                                       int chunk, int window, long long chunks, R* __restrict__ L,
                                                       R* __restrict__ U, R* __restrict__ diag, R* __restrict__ check) {
    using A = Arith<R>;
    constexpr int W = 2 * P + 1;
    constexpr int PP = P > 0 ? P : 1;
    const long long c = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c >= chunks) return;
    const long long a = first_row + c * chunk, b = (a + chunk < n) ? a + chunk : n;
    const long long start = a - window > first_row ? a - window : first_row;
    const long long stop = (b + kCheckRows < n) ? b + kCheckRows : n;   // runs into the successor's rows for the seam check
    R win[P + 1][W];
#pragma unroll
    for (int r = 0; r <= P; ++r)
#pragma unroll
        for (int e = 0; e < W; ++e) win[r][e] = (start + r < n) ? band[(start + r) * W + e] : R(0);
    for (long long k = start; k < stop; ++k) {
        // row k is final: every pivot before it has been applied
        if (k >= a) {
            if (k < b) {
                diag[k] = win[0][P];
#pragma unroll
                for (int m = 0; m < P; ++m) {
                    L[k * PP + m] = (k - P + m >= 0) ? win[0][m] : R(0);
                    U[k * PP + m] = (k + 1 + m < n) ? win[0][P + 1 + m] : R(0);
                }
            } else {
                R* chk = check + (c * kCheckRows + (k - b)) * W;
#pragma unroll
                for (int e = 0; e < W; ++e) chk[e] = win[0][e];
            }
        }
        // step(k): rows k+1 .. k+P (BandFactor::step)
        const R piv = win[0][P];
#pragma unroll
        for (int r = 1; r <= P; ++r) {
            if (k + r < n) {
                // A(k+r, k) sits at column offset P - r of row k+r
                const R l = div_rn_<R>(win[r][P - r], piv);
                win[r][P - r] = l;
#pragma unroll
                for (int j = 1; j <= P; ++j)   // A(k+r, k+j) -= l * A(k, k+j)
                    win[r][P - r + j] = A::sub(win[r][P - r + j], A::mul(l, win[0][P + j]));
            }
        }
        // slide: row k leaves, row k+P+1 enters
#pragma unroll
        for (int r = 0; r < P; ++r)
#pragma unroll
            for (int e = 0; e < W; ++e) win[r][e] = win[r + 1][e];
        const long long nr = k + P + 1;
#pragma unroll
        for (int e = 0; e < W; ++e) win[P][e] = (nr < n) ? band[nr * W + e] : R(0);
    }
}

// rows b .. b+kCheckRows-1 as chunk c computed them against chunk c+1's own output
template <typename R, int P>
__global__ void __launch_bounds__(128) check_overlap_kernel(const R* __restrict__ check, const R* __restrict__ L,
                                                            const R* __restrict__ U, const R* __restrict__ diag,
                                                            long long n, long long first_row, int chunk, long long chunks,
                                                            int* __restrict__ flag) {
    constexpr int W = 2 * P + 1;
    constexpr int PP = P > 0 ? P : 1;
    const long long v = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (v >= (chunks - 1) * kCheckRows) return;
    const long long c = v / kCheckRows;
    const int e = static_cast<int>(v - c * kCheckRows);
    const long long k = first_row + (c + 1) * chunk + e;
    if (k >= n) return;
    const R* chk = check + (c * kCheckRows + e) * W;
    bool same = chk[P] == diag[k];
#pragma unroll
    for (int m = 0; m < P; ++m) {
        if (k - P + m >= 0) same = same && chk[m] == L[k * PP + m];
        if (k + 1 + m < n) same = same && chk[P + 1 + m] == U[k * PP + m];
    }
    if (!same) atomicExch(flag, 1);
}

}  // namespace

template <typename R>
cudaError_t launch_device_band_assemble(int order, int periodic, long long n, long long K, const R* coords, const R* knots,
                                        R* band, cudaStream_t s) {
    const int bw = periodic ? order / 2 : order - 1;
    if (order < 1 || order > kFactorMaxOrder || n < 2) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(band, 0, sizeof(R) * static_cast<size_t>(n) * (2 * bw + 1), s);
    if (e != cudaSuccess) return e;
    assemble_rows_kernel<R><<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(coords, knots, order, n, K, bw, periodic, band);
    count_launch();
    return cudaGetLastError();
}

template <typename R>
cudaError_t launch_device_band_factor(int bw, long long n, long long first_row, const R* band, R* check, R* L, R* U,
                                      R* diag, int* flag, int chunk, int window, cudaStream_t s) {
    if (bw < 0 || bw > 4 || n < 2 || first_row < 0 || first_row >= n) return cudaErrorInvalidValue;
    const long long chunks = (n - first_row + chunk - 1) / chunk;
    cudaError_t e = cudaMemsetAsync(flag, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const unsigned g1 = static_cast<unsigned>((chunks + 127) / 128);
    const unsigned g2 = static_cast<unsigned>(((chunks - 1) * kCheckRows + 127) / 128);
#define BSPL_LU_CASE(P_)                                                                                                \
    case P_:                                                                                                            \
        chunk_lu_kernel<R, P_><<<g1, 128, 0, s>>>(band, n, first_row, chunk, window, chunks, L, U, diag, check);         \
        if (chunks > 1)                                                                                                 \
            check_overlap_kernel<R, P_><<<g2, 128, 0, s>>>(check, L, U, diag, n, first_row, chunk, chunks, flag);        \
        break;
    switch (bw) {
        BSPL_LU_CASE(0)
        BSPL_LU_CASE(1)
        BSPL_LU_CASE(2)
        BSPL_LU_CASE(3)
        BSPL_LU_CASE(4)
        default: return cudaErrorInvalidValue;
    }
#undef BSPL_LU_CASE
    count_launch(chunks > 1 ? 2 : 1);
    return cudaGetLastError();
}

size_t device_band_factor_check_elems(long long n, int bw, int chunk) {
    const long long chunks = (n + chunk - 1) / chunk;
    return static_cast<size_t>(chunks) * kCheckRows * (2 * bw + 1);
}

template cudaError_t launch_device_band_assemble<double>(int, int, long long, long long, const double*, const double*, double*,
                                                         cudaStream_t);
template cudaError_t launch_device_band_assemble<float>(int, int, long long, long long, const float*, const float*, float*,
                                                        cudaStream_t);
template cudaError_t launch_device_band_factor<double>(int, long long, long long, const double*, double*, double*, double*,
                                                       double*, int*, int, int, cudaStream_t);
template cudaError_t launch_device_band_factor<float>(int, long long, long long, const float*, float*, float*, float*, float*,
                                                      int*, int, int, cudaStream_t);

}  // namespace bspl
