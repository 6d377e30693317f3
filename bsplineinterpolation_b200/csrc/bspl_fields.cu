// Many fields on one small 2-D mesh, one query set for all of them (BASELINE cfg5; the
// reference's InterpolationFunctionTemplate + eval_proxy use case, InterpolationTemplate.hpp
// :118-176): each thread builds the weights of its queries ONCE, keeps them in registers, and the
// CTA then streams the fields through shared memory -- one bulk asynchronous copy (cp.async.bulk,
// SASS UBLKCP) per field, every query evaluated out of shared memory with a 16-term (cubic)
// gather-FMA.  The evaluation is bound by shared-memory wavefronts (random stencil addresses collide
// in the banks), so the CTA first deals its queries to the lanes by bank class: the lanes that one
// wavefront serves (16 for 8-byte, 32 for 4-byte elements) read stencils whose first elements lie
// in distinct banks, and so does every other stencil element, which sits at the same offset from it.
// Results return to query order through a small shared-memory stage and leave coalesced,
// [field][query]; the next field's copy is in flight while they are written out.
#include "bspl_kernels.h"

namespace bspl {

namespace {

constexpr int kFieldThreads = 512;
constexpr int kSMs = 148;       // B200
// Query slots per thread: the 2 * (O+1) weights of every slot stay in registers for the whole sweep over
// the fields, so the count follows a 96-register budget (cubic fp64: 6 slots, of which 80 % are filled).
template <typename R, int O>
constexpr int queries_per_thread() {
    constexpr int per_query = 2 * (O + 1) * static_cast<int>(sizeof(R)) / 4;
    constexpr int k = 96 / per_query;
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
    int per_cta;     // queries per CTA (a multiple of 32): CTA b owns [b * per_cta, (b + 1) * per_cta)
    int stage_off;   // byte offsets into dynamic shared memory: results of one field in query order ...
    int assign_off;  // ... and the query dealt to every (slot row, thread)
    int deriv[2];
};

constexpr unsigned short kNoQuery = 0xFFFF;
constexpr uint32_t kEmptySlot = 0xFFFFFFFFu;

template <typename R, int O>
__global__ void __launch_bounds__(kFieldThreads, 1) eval_fields_smem_kernel(const FieldsParams<R> p) {
    constexpr int W = O + 1, K = queries_per_thread<R, O>();
    constexpr int WIN = 2 * O > 0 ? 2 * O : 1;
    constexpr int NC = 128 / static_cast<int>(sizeof(R));  // lanes per shared-memory wavefront = bank classes
    constexpr int GPB = kFieldThreads / NC;                // lane groups per slot row
    constexpr int NG = K * GPB;                            // lane groups = slots of one class
    constexpr int NSLOT = K * kFieldThreads;
    extern __shared__ __align__(128) unsigned char fields_raw[];
    R* fld = reinterpret_cast<R*>(fields_raw);
    R* stage = reinterpret_cast<R*>(fields_raw + p.stage_off);
    unsigned short* assign = reinterpret_cast<unsigned short*>(fields_raw + p.assign_off);
    __shared__ uint64_t bar;
    __shared__ int taken[NC];
    const int tid = threadIdx.x;
    const uint32_t bytes = static_cast<uint32_t>(p.field_stride * sizeof(R));
    auto load_field = [&](int f) {  // thread 0, after a CTA barrier: nobody reads the buffer any more
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(bytes)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_addr(fld)),
            "l"(p.coef + static_cast<long long>(f) * p.field_stride), "r"(bytes), "r"(smem_addr(&bar))
            : "memory");
    };
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int s = tid; s < NSLOT; s += kFieldThreads) assign[s] = kNoQuery;
    if (tid < NC) taken[tid] = 0;
    __syncthreads();
    if (tid == 0 && p.n_fields > 0) load_field(0);  // under way during the set-up below

    const long long q_begin = static_cast<long long>(blockIdx.x) * p.per_cta;
    const int nq = static_cast<int>((q_begin + p.per_cta < p.q ? q_begin + p.per_cta : p.q) - q_begin);

    // deal the queries: query r goes to a lane whose index within its group equals the bank class of
    // the query's first stencil element; a full class spills into the next one (any slot is correct,
    // a matching one is conflict-free).  nq <= NSLOT, so every query finds a slot.
    for (int r = tid; r < nq; r += kFieldThreads) {
        int first = 0;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            R x = p.pts[(q_begin + r) * 2 + d];
            first += (locate<R, O>(p.ax[d], x) - O) * static_cast<int>(p.ax[d].stride);
        }
        const int cls = first & (NC - 1);
        for (int t = 0; t < NC; ++t) {
            const int c = (cls + t) & (NC - 1);
            const int m = atomicAdd(&taken[c], 1);
            if (m < NG) {
                assign[(m / GPB) * kFieldThreads + (m % GPB) * NC + c] = static_cast<unsigned short>(r);
                break;
            }
        }
    }
    __syncthreads();

    // weights of this thread's slots, kept in registers for the whole sweep over the fields
    R w0[K][W], w1[K][W];
    uint32_t slot[K];  // (query within the CTA) << 16 | first stencil element (< 2^16: the field fits 192 KB)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const unsigned short r = assign[k * kFieldThreads + tid];
        slot[k] = kEmptySlot;
#pragma unroll
        for (int i = 0; i < W; ++i) { w0[k][i] = R(0); w1[k][i] = R(0); }
        if (r != kNoQuery) {
            const long long q = q_begin + r;
            int first = 0;
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                R x = p.pts[q * 2 + d];
                const int span = locate<R, O>(p.ax[d], x);
                R tk[WIN];
                load_knot_window<R, O>(p.ax[d], span, tk);
                R* w = d == 0 ? w0[k] : w1[k];
                if (p.deriv[d] == 0) basis_funs<R, O>(tk, x, O, w);
                else deriv_weights<R, O>(tk, x, p.deriv[d], w);
                first += (span - O) * static_cast<int>(p.ax[d].stride);
            }
            slot[k] = (static_cast<uint32_t>(r) << 16) | static_cast<uint32_t>(first);
        }
    }
    const int s0 = static_cast<int>(p.ax[0].stride);
    uint32_t parity = 0;
    for (int f = 0; f < p.n_fields; ++f) {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done)
                : "r"(smem_addr(&bar)), "r"(parity)
                : "memory");
        }
        parity ^= 1;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (slot[k] != kEmptySlot) {
                const R* c = fld + (slot[k] & 0xFFFFu);
                R v = R(0);
#pragma unroll
                for (int i = 0; i < W; ++i) {
                    R a = R(0);
#pragma unroll
                    for (int j = 0; j < W; ++j) a += c[i * s0 + j] * w1[k][j];
                    v += a * w0[k][i];
                }
                stage[slot[k] >> 16] = v;
            }
        }
        __syncthreads();  // the field has been consumed, the stage is complete
        if (tid == 0 && f + 1 < p.n_fields) load_field(f + 1);
        R* o = p.out + static_cast<long long>(f) * p.q + q_begin;
        for (int r = tid; r < nq; r += kFieldThreads) o[r] = stage[r];
        __syncthreads();  // the stage is free for the next field
    }
}

template <typename R, int O>
cudaError_t fields_O(const EvalArgs<R>& a, cudaStream_t s) {
    FieldsParams<R> p;
    for (int d = 0; d < 2; ++d) { p.ax[d] = a.ax[d]; p.deriv[d] = a.deriv[d]; }
    p.coef = a.coef; p.field_stride = a.field_stride; p.n_fields = a.n_fields;
    p.pts = a.pts; p.out = a.out; p.q = a.q;
    auto k = eval_fields_smem_kernel<R, O>;
    // One CTA per SM at a time (the field fills shared memory), every CTA sweeps all fields: the batch is
    // cut into a whole number of waves of equal CTAs instead of full CTAs plus a ragged last wave
    // (2^20 queries: 444 CTAs of 2 368 queries in 3 waves).  A CTA is filled to at most 80 % of its
    // slots, so that nearly every query finds a lane of its own bank class.
    const long long slots = static_cast<long long>(kFieldThreads) * queries_per_thread<R, O>();
    const long long cap = slots * 4 / 5 / 32 * 32;
    const long long waves = (a.q + cap * kSMs - 1) / (cap * kSMs);
    long long grid = waves * kSMs;
    long long per_cta = ((a.q + grid - 1) / grid + 31) / 32 * 32;
    if (per_cta > cap) per_cta = cap;
    grid = (a.q + per_cta - 1) / per_cta;
    p.per_cta = static_cast<int>(per_cta);
    const long long field_bytes = (a.field_stride * static_cast<long long>(sizeof(R)) + 127) / 128 * 128;
    const long long stage_bytes = (per_cta * static_cast<long long>(sizeof(R)) + 127) / 128 * 128;
    p.stage_off = static_cast<int>(field_bytes);
    p.assign_off = static_cast<int>(field_bytes + stage_bytes);
    const int smem = static_cast<int>(field_bytes + stage_bytes + slots * 2);
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    k<<<static_cast<unsigned>(grid), kFieldThreads, smem, s>>>(p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace

template <typename R>
bool fields_smem_eligible(const EvalArgs<R>& a) {
    const long long bytes = a.field_stride * static_cast<long long>(sizeof(R));
    return a.dim == 2 && a.order <= 5 && a.mode == kValue && a.n_fields >= 8 && a.q >= 4096 && bytes % 16 == 0 &&
           bytes <= 192 * 1024 && a.field_stride <= 65535;  // + 31 KB of stage and slot table; 16-bit element offsets
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
