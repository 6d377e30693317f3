// Many fields, one query set, as a contraction (BASELINE cfg5; the reference's
// InterpolationFunctionTemplate + eval_proxy use case, InterpolationTemplate.hpp:118-176,
// BSpline.hpp:244-297): with the queries sorted by cell, all queries of a cell read the same
// K = (O+1)^D control points of every field, so
//     out[q, f] = sum_k W[q, k] * C[cell(q) + off_k, f]
// is a small dense product per cell: (queries of the cell) x K times K x (fields).  The control
// points are kept field-minor (one transposed copy per function, made on first use), so that a
// warp reads C[., f0 .. f0+127] and writes out[q, f0 .. f0+127] as whole lines.
//
//   1. fields_key_kernel      locate each query, key = element offset of its cell, histogram
//   2. fields_plan_kernel     exclusive scan of the histogram, work list (key, begin, end) in
//                             chunks of at most kQueryChunk queries
//   3. fields_weights_kernel  counting-sort scatter: original index and the K tensor-product
//                             weights of every query, in cell order
//   4. fields_contract_kernel persistent CTAs pull (work item, field block) units: the cell's
//                             K x TF control points of the thread's fields sit in registers, the
//                             weights of the item's queries in shared memory (every lane reads the
//                             same weight: one broadcast wavefront per load), TQ x TF accumulators
//                             per thread, results stored as TF-wide vectors, query-major.
//
// Results are query-major, out[q][n_fields] -- the layout the reference itself produces for a
// vector-valued T (one T{...} per query, interpolation-test.cpp:674-703).  A field-major result
// with the queries of a cell scattered over 8-byte slots would be bound by L2 transactions, so
// the field-major entry point transposes blocks of this kernel's output instead (bspl_capi.cu).
#include <type_traits>

#include "bspl_kernels.h"

namespace bspl {

namespace {

constexpr int kSMs = 148;
constexpr int kContractThreads = 128;
constexpr int kMaxQueryChunk = 128;    // queries per work item ...
constexpr int kWeightSlots = 2048;     // ... whose K weights each fit this many shared-memory elements

template <typename R>
struct FieldsSortParams {
    int dim;
    AxisParams<R> ax[kMaxDim];
    int deriv[kMaxDim];
    const R* pts;
    long long q;
};

template <typename R, int O>
__device__ __forceinline__ uint32_t cell_key(const FieldsSortParams<R>& p, long long q) {
    long long key = 0;
    for (int d = 0; d < p.dim; ++d) {
        R x = p.pts[q * p.dim + d];
        key += static_cast<long long>(locate<R, O>(p.ax[d], x) - O) * p.ax[d].stride;
    }
    return static_cast<uint32_t>(key);
}

template <typename R, int O>
__global__ void __launch_bounds__(256) fields_key_kernel(const FieldsSortParams<R> p, uint32_t* __restrict__ counts) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < p.q; q += stride)
        atomicAdd(&counts[cell_key<R, O>(p, q)], 1u);
}

// Single CTA: cursor[key] = exclusive scan of counts; work[3 i + {0,1,2}] = {key, begin, end}.
__global__ void __launch_bounds__(1024) fields_plan_kernel(const uint32_t* __restrict__ counts, int n_keys, int kQueryChunk,
                                                           uint32_t* __restrict__ cursor, uint32_t* __restrict__ work,
                                                           uint32_t* __restrict__ n_work, uint32_t* __restrict__ next_unit) {
    __shared__ uint32_t part_q[1024];
    __shared__ uint32_t part_w[1024];
    const int t = threadIdx.x;
    const int per = (n_keys + 1023) / 1024;
    const int b = min(n_keys, t * per), e = min(n_keys, b + per);
    uint32_t sq = 0, sw = 0;
    for (int i = b; i < e; ++i) {
        const uint32_t c = counts[i];
        sq += c;
        sw += (c + kQueryChunk - 1) / kQueryChunk;
    }
    part_q[t] = sq;
    part_w[t] = sw;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        uint32_t vq = 0, vw = 0;
        if (t >= off) { vq = part_q[t - off]; vw = part_w[t - off]; }
        __syncthreads();
        part_q[t] += vq;
        part_w[t] += vw;
        __syncthreads();
    }
    uint32_t oq = part_q[t] - sq, ow = part_w[t] - sw;
    for (int i = b; i < e; ++i) {
        const uint32_t c = counts[i];
        cursor[i] = oq;
        for (uint32_t s = 0; s < c; s += kQueryChunk) {
            work[3 * ow + 0] = static_cast<uint32_t>(i);
            work[3 * ow + 1] = oq + s;
            work[3 * ow + 2] = oq + min(c, s + static_cast<uint32_t>(kQueryChunk));
            ++ow;
        }
        oq += c;
    }
    if (t == 1023) { *n_work = part_w[1023]; *next_unit = 0; }
}

// K = (O+1)^dim tensor-product weights, last axis fastest (the order of the offset table)
template <typename R, int O>
__global__ void __launch_bounds__(256) fields_weights_kernel(const FieldsSortParams<R> p, uint32_t* __restrict__ cursor,
                                                             uint32_t* __restrict__ idx_sorted, R* __restrict__ w_sorted,
                                                             int K) {
    constexpr int W = O + 1;
    constexpr int WIN = 2 * O > 0 ? 2 * O : 1;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < p.q; q += stride) {
        R w[kMaxDim][W];
        long long key = 0;
        for (int d = 0; d < p.dim; ++d) {
            R x = p.pts[q * p.dim + d];
            const int span = locate<R, O>(p.ax[d], x);
            R tk[WIN];
            load_knot_window<R, O>(p.ax[d], span, tk);
            if (p.deriv[d] == 0) basis_funs<R, O>(tk, x, O, w[d]);
            else deriv_weights<R, O>(tk, x, p.deriv[d], w[d]);
            key += static_cast<long long>(span - O) * p.ax[d].stride;
        }
        const uint32_t pos = atomicAdd(&cursor[key], 1u);
        idx_sorted[pos] = static_cast<uint32_t>(q);
        R* dst = w_sorted + static_cast<long long>(pos) * K;
        if (p.dim == 1) {
#pragma unroll
            for (int i = 0; i < W; ++i) dst[i] = w[0][i];
        } else if (p.dim == 2) {
#pragma unroll
            for (int i = 0; i < W; ++i)
#pragma unroll
                for (int j = 0; j < W; ++j) dst[i * W + j] = w[0][i] * w[1][j];
        } else {
#pragma unroll
            for (int i = 0; i < W; ++i)
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const R wij = w[0][i] * w[1][j];
#pragma unroll
                    for (int k = 0; k < W; ++k) dst[(i * W + j) * W + k] = wij * w[2][k];
                }
        }
    }
}

template <typename R, int TF> struct FieldVec;
template <> struct __align__(32) FieldVec<double, 4> { double v[4]; };
template <> struct __align__(16) FieldVec<double, 2> { double v[2]; };
template <> struct __align__(8) FieldVec<double, 1> { double v[1]; };
template <> struct __align__(16) FieldVec<float, 4> { float v[4]; };
template <> struct __align__(8) FieldVec<float, 2> { float v[2]; };
template <> struct __align__(4) FieldVec<float, 1> { float v[1]; };

template <typename R>
struct ContractParams {
    const R* coef_t;            // [field_stride][n_fields] field-minor control points
    const R* w_sorted;          // [q][K]
    const uint32_t* idx_sorted; // [q]
    const uint32_t* work;
    const uint32_t* n_work;
    uint32_t* next_unit;
    R* out;                     // element (query, field) at query * out_stride + field - field_begin
    long long out_stride;
    int n_fields;
    int field_begin, field_end; // fields evaluated by this launch
    int n_fb;                   // field blocks of kContractThreads * TF
    int off[64];                // element offset of stencil term k from the cell's first control point
};

template <typename R, int K, int TF>
__global__ void __launch_bounds__(kContractThreads) fields_contract_kernel(const ContractParams<R> p) {
    constexpr int TQ = 4;
    constexpr int FB = kContractThreads * TF;
    using Vec = FieldVec<R, TF>;
    __shared__ R s_w[kWeightSlots];
    __shared__ uint32_t s_idx[kMaxQueryChunk];
    __shared__ uint32_t s_unit;
    const int tid = threadIdx.x;
    const uint32_t n_units = *p.n_work * static_cast<uint32_t>(p.n_fb);
    for (;;) {
        if (tid == 0) s_unit = atomicAdd(p.next_unit, 1u);
        __syncthreads();  // also: every thread is past its reads of the previous unit's weights
        const uint32_t unit = s_unit;
        if (unit >= n_units) break;
        const uint32_t item = unit / p.n_fb, fb = unit - item * p.n_fb;
        const uint32_t key = p.work[3 * item], begin = p.work[3 * item + 1], end = p.work[3 * item + 2];
        const int nq = static_cast<int>(end - begin);
        // the item's weights and original indices -> shared memory (contiguous in cell order)
        for (int e = tid; e < nq * K; e += kContractThreads) s_w[e] = p.w_sorted[static_cast<long long>(begin) * K + e];
        for (int e = tid; e < nq; e += kContractThreads) s_idx[e] = p.idx_sorted[begin + e];
        // this thread's TF fields of the cell's K control points -> registers
        const int f0 = p.field_begin + static_cast<int>(fb) * FB + tid * TF;
        const bool active = f0 < p.field_end;
        Vec c[K];
        if (active) {
            const R* base = p.coef_t + static_cast<long long>(key) * p.n_fields + f0;
#pragma unroll
            for (int k = 0; k < K; ++k)
                c[k] = *reinterpret_cast<const Vec*>(base + static_cast<long long>(p.off[k]) * p.n_fields);
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int f = 0; f < TF; ++f) c[k].v[f] = R(0);
        }
        __syncthreads();
        if (active) {
            R* obase = p.out + (f0 - p.field_begin);
            int q0 = 0;
            for (; q0 + TQ <= nq; q0 += TQ) {
                R acc[TQ][TF];
#pragma unroll
                for (int j = 0; j < TQ; ++j)
#pragma unroll
                    for (int f = 0; f < TF; ++f) acc[j][f] = R(0);
#pragma unroll
                for (int k = 0; k < K; ++k) {
#pragma unroll
                    for (int j = 0; j < TQ; ++j) {
                        const R w = s_w[(q0 + j) * K + k];
#pragma unroll
                        for (int f = 0; f < TF; ++f) acc[j][f] = fma(c[k].v[f], w, acc[j][f]);
                    }
                }
#pragma unroll
                for (int j = 0; j < TQ; ++j) {
                    Vec r;
#pragma unroll
                    for (int f = 0; f < TF; ++f) r.v[f] = acc[j][f];
                    *reinterpret_cast<Vec*>(obase + static_cast<long long>(s_idx[q0 + j]) * p.out_stride) = r;
                }
            }
            for (; q0 < nq; ++q0) {
                R acc[TF];
#pragma unroll
                for (int f = 0; f < TF; ++f) acc[f] = R(0);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const R w = s_w[q0 * K + k];
#pragma unroll
                    for (int f = 0; f < TF; ++f) acc[f] = fma(c[k].v[f], w, acc[f]);
                }
                Vec r;
#pragma unroll
                for (int f = 0; f < TF; ++f) r.v[f] = acc[f];
                *reinterpret_cast<Vec*>(obase + static_cast<long long>(s_idx[q0]) * p.out_stride) = r;
            }
        }
    }
}

// registers hold K * TF control points per thread: wider field vectors for small stencils
template <typename R, int K>
constexpr int fields_per_thread() {
    return sizeof(R) == 4 ? (K <= 36 ? 4 : 2) : (K <= 16 ? 4 : (K <= 36 ? 2 : 1));
}

template <typename R, int K>
cudaError_t contract_K(const FieldsContractArgs<R>& a, const FieldsScratch& sc, cudaStream_t s) {
    constexpr int TF = fields_per_thread<R, K>();
    if (a.n_fields % TF != 0 || a.field_begin % TF != 0 || a.out_stride % TF != 0 ||
        (reinterpret_cast<uintptr_t>(a.out) % (TF * sizeof(R))) != 0 ||
        (reinterpret_cast<uintptr_t>(a.coef_t) % (TF * sizeof(R))) != 0)
        return cudaErrorNotSupported;
    ContractParams<R> p;
    p.coef_t = a.coef_t; p.w_sorted = static_cast<const R*>(sc.w_sorted); p.idx_sorted = sc.idx_sorted;
    p.work = sc.work; p.n_work = sc.n_work; p.next_unit = sc.next_unit;
    p.out = a.out; p.out_stride = a.out_stride; p.n_fields = a.n_fields;
    p.field_begin = a.field_begin; p.field_end = a.field_end;
    constexpr int FB = kContractThreads * TF;
    p.n_fb = (a.field_end - a.field_begin + FB - 1) / FB;
    for (int k = 0; k < K; ++k) p.off[k] = a.off[k];
    cudaError_t e = cudaMemsetAsync(sc.next_unit, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fields_contract_kernel<R, K, TF>, kContractThreads, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    fields_contract_kernel<R, K, TF><<<kSMs * per_sm, kContractThreads, 0, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

template <typename R, int O>
cudaError_t sort_O(const FieldsContractArgs<R>& a, const FieldsScratch& sc, cudaStream_t s) {
    FieldsSortParams<R> p;
    p.dim = a.dim;
    for (int d = 0; d < a.dim; ++d) { p.ax[d] = a.ax[d]; p.deriv[d] = a.deriv[d]; }
    p.pts = a.pts; p.q = a.q;
    cudaError_t e = cudaMemsetAsync(sc.counts, 0, sizeof(uint32_t) * static_cast<size_t>(a.n_keys), s);
    if (e != cudaSuccess) return e;
    const int grid = static_cast<int>(std::min<long long>((a.q + 255) / 256, kSMs * 8));
    fields_key_kernel<R, O><<<grid, 256, 0, s>>>(p, sc.counts);
    fields_plan_kernel<<<1, 1024, 0, s>>>(sc.counts, a.n_keys, a.chunk, sc.cursor, sc.work, sc.n_work, sc.next_unit);
    fields_weights_kernel<R, O><<<grid, 256, 0, s>>>(p, sc.cursor, sc.idx_sorted, static_cast<R*>(sc.w_sorted), a.K);
    count_launch(3);
    return cudaGetLastError();
}

}  // namespace

int fields_query_chunk(int K) {
    int c = kWeightSlots / (K > 0 ? K : 1);
    if (c > kMaxQueryChunk) c = kMaxQueryChunk;
    c = c / 4 * 4;
    return c < 4 ? 4 : c;
}

size_t fields_scratch_bytes(long long q, int n_keys, int K, size_t elem, size_t* offsets) {
    const int kQueryChunk = fields_query_chunk(K);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    offsets[0] = take(sizeof(uint32_t) * static_cast<size_t>(n_keys));              // counts
    offsets[1] = take(sizeof(uint32_t) * static_cast<size_t>(n_keys));              // cursor
    offsets[2] = take(sizeof(uint32_t) * static_cast<size_t>(q));                   // idx_sorted
    offsets[3] = take(elem * static_cast<size_t>(q) * K);                           // w_sorted
    offsets[4] = take(sizeof(uint32_t) * 3 * (static_cast<size_t>(q) / kQueryChunk + n_keys + 1));  // work
    offsets[5] = take(256);                                                         // n_work, next_unit
    return off;
}

FieldsScratch fields_scratch_view(void* base, long long q, int n_keys, int K, size_t elem) {
    size_t off[6];
    fields_scratch_bytes(q, n_keys, K, elem, off);
    unsigned char* b = static_cast<unsigned char*>(base);
    FieldsScratch sc;
    sc.counts = reinterpret_cast<uint32_t*>(b + off[0]);
    sc.cursor = reinterpret_cast<uint32_t*>(b + off[1]);
    sc.idx_sorted = reinterpret_cast<uint32_t*>(b + off[2]);
    sc.w_sorted = b + off[3];
    sc.work = reinterpret_cast<uint32_t*>(b + off[4]);
    sc.n_work = reinterpret_cast<uint32_t*>(b + off[5]);
    sc.next_unit = sc.n_work + 1;
    return sc;
}

bool fields_contract_supported(int dim, int order) {
    if (dim > 3 || order > 5) return false;   // the sort and weight kernels are instantiated for those
    int K = 1;
    for (int d = 0; d < dim; ++d) K *= order + 1;
    switch (K) {
        case 1: case 2: case 3: case 4: case 5: case 6: case 8: case 9: case 16: case 25: case 27: case 36: case 64:
            return true;
        default: return false;
    }
}

template <typename R>
cudaError_t launch_fields_sort(const FieldsContractArgs<R>& a, const FieldsScratch& sc, cudaStream_t s) {
    if (a.q <= 0) return cudaSuccess;
    if (a.q >= (1ll << 32)) return cudaErrorInvalidValue;
    switch (a.order) {
        case 0: return sort_O<R, 0>(a, sc, s);
        case 1: return sort_O<R, 1>(a, sc, s);
        case 2: return sort_O<R, 2>(a, sc, s);
        case 3: return sort_O<R, 3>(a, sc, s);
        case 4: return sort_O<R, 4>(a, sc, s);
        case 5: return sort_O<R, 5>(a, sc, s);
        default: return cudaErrorInvalidValue;
    }
}

template <typename R>
cudaError_t launch_fields_contract(const FieldsContractArgs<R>& a, const FieldsScratch& sc, cudaStream_t s) {
    if (a.q <= 0 || a.field_end <= a.field_begin) return cudaSuccess;
    switch (a.K) {
#define BSPL_CONTRACT_CASE(K_) case K_: return contract_K<R, K_>(a, sc, s);
        BSPL_CONTRACT_CASE(1)
        BSPL_CONTRACT_CASE(2)
        BSPL_CONTRACT_CASE(3)
        BSPL_CONTRACT_CASE(4)
        BSPL_CONTRACT_CASE(5)
        BSPL_CONTRACT_CASE(6)
        BSPL_CONTRACT_CASE(8)
        BSPL_CONTRACT_CASE(9)
        BSPL_CONTRACT_CASE(16)
        BSPL_CONTRACT_CASE(25)
        BSPL_CONTRACT_CASE(27)
        BSPL_CONTRACT_CASE(36)
        BSPL_CONTRACT_CASE(64)
#undef BSPL_CONTRACT_CASE
        default: return cudaErrorNotSupported;
    }
}

template cudaError_t launch_fields_sort<double>(const FieldsContractArgs<double>&, const FieldsScratch&, cudaStream_t);
template cudaError_t launch_fields_sort<float>(const FieldsContractArgs<float>&, const FieldsScratch&, cudaStream_t);
template cudaError_t launch_fields_contract<double>(const FieldsContractArgs<double>&, const FieldsScratch&, cudaStream_t);
template cudaError_t launch_fields_contract<float>(const FieldsContractArgs<float>&, const FieldsScratch&, cudaStream_t);

}  // namespace bspl
