// Many fields on one small 2-D mesh, one query set for all of them (BASELINE cfg5; the
// reference's InterpolationFunctionTemplate + eval_proxy use case, InterpolationTemplate.hpp
// :118-176): each thread locates its queries and builds their weights ONCE, keeps them in
// registers, and the CTA then streams the fields through shared memory -- one bulk asynchronous
// copy (cp.async.bulk, SASS UBLKCP) per field, every query evaluated out of shared memory with a
// 16-term (cubic) gather-FMA.  Output is [field][query], written coalesced.
#include "bspl_kernels.h"

namespace bspl {

namespace {

constexpr int kFieldThreads = 512;
constexpr int kSMs = 148;       // B200
// Queries per thread: their 2 * (O+1) weights stay in registers for the whole sweep over the fields, so
// the count follows an 80-register budget (cubic fp64: 5; ptxas: 128 registers, no spills).
template <typename R, int O>
constexpr int queries_per_thread() {
    constexpr int per_query = 2 * (O + 1) * static_cast<int>(sizeof(R)) / 4;
    constexpr int k = 80 / per_query;
    return k > 6 ? 6 : (k < 2 ? 2 : k);
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

template <typename R>
struct FieldsParams {
    AxisParams<R> ax[2];
    const R* coef;
    long long field_stride;  // elements; the padded field is contiguous
    int n_fields;
    const R* pts;
    R* out;
    long long q;
    int per_cta;  // queries per CTA (a multiple of 32): CTA b owns [b * per_cta, (b + 1) * per_cta)
    int deriv[2];
};

template <typename R, int O>
__global__ void __launch_bounds__(kFieldThreads, 1) eval_fields_smem_kernel(const FieldsParams<R> p) {
    constexpr int W = O + 1, K = queries_per_thread<R, O>();
    constexpr int WIN = 2 * O > 0 ? 2 * O : 1;
    extern __shared__ __align__(128) unsigned char fields_raw[];
    R* fld = reinterpret_cast<R*>(fields_raw);
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // this thread's queries: q0 + k * kFieldThreads, so that a warp's stores are contiguous
    const long long q_begin = static_cast<long long>(blockIdx.x) * p.per_cta;
    const long long q_end = q_begin + p.per_cta < p.q ? q_begin + p.per_cta : p.q;
    const long long q0 = q_begin + tid;
    R w0[K][W], w1[K][W];
    int off[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const long long q = q0 + static_cast<long long>(k) * kFieldThreads;
        off[k] = 0;
#pragma unroll
        for (int i = 0; i < W; ++i) { w0[k][i] = R(0); w1[k][i] = R(0); }
        if (q < q_end) {
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                R x = p.pts[q * 2 + d];
                const int span = locate<R, O>(p.ax[d], x);
                R tk[WIN];
                load_knot_window<R, O>(p.ax[d], span, tk);
                R* w = d == 0 ? w0[k] : w1[k];
                if (p.deriv[d] == 0) basis_funs<R, O>(tk, x, O, w);
                else deriv_weights<R, O>(tk, x, p.deriv[d], w);
                off[k] += (span - O) * static_cast<int>(p.ax[d].stride);
            }
        }
    }
    const uint32_t bytes = static_cast<uint32_t>(p.field_stride * sizeof(R));
    const int s0 = static_cast<int>(p.ax[0].stride);
    uint32_t parity = 0;
    for (int f = 0; f < p.n_fields; ++f) {
        __syncthreads();  // everyone is done with the previous field
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(bytes)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_addr(fld)),
                "l"(p.coef + static_cast<long long>(f) * p.field_stride), "r"(bytes), "r"(smem_addr(&bar))
                : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done)
                : "r"(smem_addr(&bar)), "r"(parity)
                : "memory");
        }
        parity ^= 1;
        R* o = p.out + static_cast<long long>(f) * p.q;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const long long q = q0 + static_cast<long long>(k) * kFieldThreads;
            if (q < q_end) {
                const R* c = fld + off[k];
                R v = R(0);
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    R a = R(0);
#pragma unroll
                    for (int j = 0; j < W; ++j) a += c[i * s0 + j] * w1[k][j];
                    v += a * w0[k][i];
                }
                o[q] = v;
            }
        }
    }
}

template <typename R, int O>
cudaError_t fields_O(const EvalArgs<R>& a, cudaStream_t s) {
    FieldsParams<R> p;
    for (int d = 0; d < 2; ++d) { p.ax[d] = a.ax[d]; p.deriv[d] = a.deriv[d]; }
    p.coef = a.coef; p.field_stride = a.field_stride; p.n_fields = a.n_fields;
    p.pts = a.pts; p.out = a.out; p.q = a.q;
    const int smem = static_cast<int>(a.field_stride * sizeof(R));
    auto k = eval_fields_smem_kernel<R, O>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    // One CTA per SM at a time (the field fills shared memory), every CTA sweeps all fields: the batch is
    // cut into a whole number of waves of equal CTAs instead of full CTAs plus a ragged last wave
    // (2^20 queries: 444 CTAs of 2 368 queries in 3 waves, not 512 of 2 048 in 3.46).
    const long long cap = static_cast<long long>(kFieldThreads) * queries_per_thread<R, O>();
    const long long waves = (a.q + cap * kSMs - 1) / (cap * kSMs);
    long long grid = waves * kSMs;
    long long per_cta = ((a.q + grid - 1) / grid + 31) / 32 * 32;
    if (per_cta > cap) per_cta = cap / 32 * 32;
    grid = (a.q + per_cta - 1) / per_cta;
    p.per_cta = static_cast<int>(per_cta);
    k<<<static_cast<unsigned>(grid), kFieldThreads, smem, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

template <typename R>
bool fields_smem_eligible(const EvalArgs<R>& a) {
    const long long bytes = a.field_stride * static_cast<long long>(sizeof(R));
    return a.dim == 2 && a.mode == kValue && a.n_fields >= 8 && a.q >= 4096 && bytes % 16 == 0 &&
           bytes <= 200 * 1024 && a.ax[0].stride < (1 << 20);
}

template <typename R>
cudaError_t launch_eval_fields_smem(const EvalArgs<R>& a, cudaStream_t s) {
    switch (a.order) {
        case 0: return fields_O<R, 0>(a, s);
        case 1: return fields_O<R, 1>(a, s);
        case 2: return fields_O<R, 2>(a, s);
        case 3: return fields_O<R, 3>(a, s);
        case 4: return fields_O<R, 4>(a, s);
        case 5: return fields_O<R, 5>(a, s);
        default: return cudaErrorInvalidValue;
    }
}

template bool fields_smem_eligible<double>(const EvalArgs<double>&);
template bool fields_smem_eligible<float>(const EvalArgs<float>&);
template cudaError_t launch_eval_fields_smem<double>(const EvalArgs<double>&, cudaStream_t);
template cudaError_t launch_eval_fields_smem<float>(const EvalArgs<float>&, cudaStream_t);

}  // namespace bspl
