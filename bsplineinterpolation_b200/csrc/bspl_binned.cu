// Cell-binned evaluation (3-D): queries are grouped by coefficient tile, each
// tile's (T+O)^3 brick is staged into shared memory with one TMA box copy
// (cp.async.bulk.tensor.3d) and every query of the tile is evaluated out of
// shared memory.  This replaces the 16x-inflated INTP_CELL_LAYOUT array of the
// reference (BSpline.hpp:677-758) -- whose purpose is to make a query's stencil
// a contiguous read -- with the device analogue: a compact, ghost-padded array
// read once per tile, plus a sorted copy of the queries (coordinates + original index).
//
// Pipeline, all on the caller's stream, no host synchronisation:
//   1. key_count    : locate each query, key = (tile, x, y cell in tile), histogram over keys
//   2. plan         : exclusive scan of the histogram, cursors, per-tile work list
//                     (tile, begin, end) in chunks of kChunk queries
//   3. scatter      : rec[cursor[key]++] = {x, y, z, q}  (counting sort, tile-major)
//   4. eval         : persistent CTAs pull work items; TMA brick -> smem;
//                     gather-FMA from smem; results written back to out[q]
#include <cuda.h>

#include "bspl_kernels.h"
#include "bspl_tma.cuh"

namespace bspl {

namespace {

constexpr int kSMs = 148;
constexpr int kChunk = 4096;        // queries per work item
constexpr int kEvalThreads = 256;

// Odd orders keep two copies of the brick, the second shifted by one element along z, so
// that a stencil row is always a pair-aligned vector read (17^2 x 18 doubles each);
// even orders (odd stencil width) use one brick and scalar reads.
constexpr int tile_edge_rt(int O) { return (O % 2) ? 17 - O : (O == 0 ? 16 : 18 - O); }
template <int O> constexpr int tile_edge() { return tile_edge_rt(O); }
template <int O> constexpr bool dual_brick() { return O % 2 == 1; }
template <int O> constexpr int brick_edge() { return tile_edge<O>() + O; }
// inner box extent: rows must be a multiple of 16 bytes for TMA
template <typename R, int O> constexpr int brick_pitch() {
    return (brick_edge<O>() + int(16 / sizeof(R)) - 1) / int(16 / sizeof(R)) * int(16 / sizeof(R));
}

template <typename R>
struct BinParams {
    AxisParams<R> ax[3];
    const R* pts;
    long long q;
    int ntile[3];
    int n_tiles;
};

// Sort key of a query: (tile, x cell in tile, y cell in tile).  Queries of one key share
// their stencil's 16 (x, y) rows and differ only in the z offset inside those rows, so a
// warp of consecutive sorted queries reads shared memory without bank conflicts.
template <typename R, int O>
__device__ __forceinline__ uint32_t key_of_point(const BinParams<R>& p, long long q) {
    constexpr int T = tile_edge<O>();
    int c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) c[d] = locate_quick<R, O>(p.ax[d], p.pts[q * 3 + d]) - O;
    const int tx = c[0] / T, ty = c[1] / T, tz = c[2] / T;
    const int tile = (tx * p.ntile[1] + ty) * p.ntile[2] + tz;
    return static_cast<uint32_t>(tile) * (T * T) + (c[0] - tx * T) * T + (c[1] - ty * T);
}

template <typename R, int O>
__global__ void __launch_bounds__(512) key_count_kernel(const BinParams<R> p, uint32_t* __restrict__ counts) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < p.q; q += stride)
        atomicAdd(&counts[key_of_point<R, O>(p, q)], 1u);
}

// Per-tile totals of the key histogram: one warp per tile.
__global__ void __launch_bounds__(256) tile_totals_kernel(const uint32_t* __restrict__ counts, int n_tiles, int bpt,
                                                          uint32_t* __restrict__ totals) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_tiles) return;
    uint32_t c = 0;
    for (int k = lane; k < bpt; k += 32) c += counts[static_cast<long long>(warp) * bpt + k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
    if (lane == 0) totals[warp] = c;
}

// cursor[key] = tile offset + exclusive scan of the tile's key counts: one warp per tile.
__global__ void __launch_bounds__(256) key_cursor_kernel(const uint32_t* __restrict__ counts, int n_tiles, int bpt,
                                                         const uint32_t* __restrict__ tile_off,
                                                         uint32_t* __restrict__ cursor) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_tiles) return;
    uint32_t run = tile_off[warp];
    for (int k0 = 0; k0 < bpt; k0 += 32) {
        const int k = k0 + lane;
        const uint32_t c = k < bpt ? counts[static_cast<long long>(warp) * bpt + k] : 0;
        uint32_t incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (k < bpt) cursor[static_cast<long long>(warp) * bpt + k] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// Single CTA: exclusive scan of the per-tile totals -> tile offsets, and the work list.
// work[3*i + {0,1,2}] = {tile, begin, end}; *n_work = number of items.
__global__ void __launch_bounds__(1024) plan_kernel(const uint32_t* __restrict__ counts, int n_tiles,
                                                    uint32_t* __restrict__ cursor, uint32_t* __restrict__ work,
                                                    uint32_t* __restrict__ n_work, uint32_t* __restrict__ next_item) {
    __shared__ unsigned long long part_q[1024];
    __shared__ uint32_t part_w[1024];
    const int t = threadIdx.x;
    const int per = (n_tiles + 1023) / 1024;
    const int b = min(n_tiles, t * per), e = min(n_tiles, b + per);
    unsigned long long sq = 0;
    uint32_t sw = 0;
    for (int i = b; i < e; ++i) {
        const uint32_t c = counts[i];
        sq += c;
        sw += (c + kChunk - 1) / kChunk;
    }
    part_q[t] = sq;
    part_w[t] = sw;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partials
    for (int off = 1; off < 1024; off <<= 1) {
        unsigned long long vq = 0;
        uint32_t vw = 0;
        if (t >= off) { vq = part_q[t - off]; vw = part_w[t - off]; }
        __syncthreads();
        part_q[t] += vq;
        part_w[t] += vw;
        __syncthreads();
    }
    unsigned long long oq = part_q[t] - sq;
    uint32_t ow = part_w[t] - sw;
    for (int i = b; i < e; ++i) {
        const uint32_t c = counts[i];
        cursor[i] = static_cast<uint32_t>(oq);
        for (uint32_t s = 0; s < c; s += kChunk) {
            work[3 * ow + 0] = static_cast<uint32_t>(i);
            work[3 * ow + 1] = static_cast<uint32_t>(oq) + s;
            work[3 * ow + 2] = static_cast<uint32_t>(oq) + min(c, s + kChunk);
            ++ow;
        }
        oq += c;
    }
    if (t == 1023) { *n_work = part_w[1023]; *next_item = 0; }
}

// Sorted query record: coordinates + original index, 4 x sizeof(R) bytes, written and read
// as whole aligned sectors (the scattered side of the sort is the write, which needs no
// latency hiding; the evaluation kernel then streams its queries).
template <typename R> struct Out4;
template <> struct __align__(32) Out4<double> { double v, g0, g1, g2; };
template <> struct __align__(16) Out4<float> { float v, g0, g1, g2; };
template <typename R> struct Rec;
template <> struct __align__(32) Rec<double> { double x, y, z; unsigned long long idx; };
template <> struct __align__(16) Rec<float> { float x, y, z; uint32_t idx; };

// The key is recomputed here rather than stored by the counting pass: this kernel waits on
// random sector writes (issue slots ~2 % busy), so the arithmetic is free and 8 bytes of traffic
// per query disappear.
template <typename R, int O>
__global__ void __launch_bounds__(512) scatter_kernel(const BinParams<R> p, uint32_t* __restrict__ cursor,
                                                      Rec<R>* __restrict__ rec) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < p.q; i += stride) {
        Rec<R> r;
        r.x = p.pts[3 * i]; r.y = p.pts[3 * i + 1]; r.z = p.pts[3 * i + 2];
        r.idx = static_cast<decltype(r.idx)>(i);
        const uint32_t pos = atomicAdd(&cursor[key_of_point<R, O>(p, i)], 1u);
        rec[pos] = r;
    }
}

template <typename R>
struct BinEvalParams {
    AxisParams<R> ax[3];
    const Rec<R>* rec;
    R* out;
    const uint32_t* work;
    const uint32_t* n_work;
    uint32_t* next_item;
    int ntile[3];
    int deriv[3];
};

template <typename R, int O, bool GRAD>
__device__ __forceinline__ int axis_weights(const AxisParams<R>& a, R x, int k, R* w, R* dw) {
    constexpr int WIN = 2 * O > 0 ? 2 * O : 1;
    const int span = locate<R, O>(a, x);
    R tk[WIN];
    load_knot_window<R, O>(a, span, tk);
    if (GRAD) {
        basis_and_deriv<R, O>(tk, x, w, dw);
    } else {
        if (k == 0) basis_funs<R, O>(tk, x, O, w);
        else deriv_weights<R, O>(tk, x, k, w);
    }
    return span - O;
}

// Same, for tiles whose every stencil stays clear of the clamped end knots of a uniform
// axis (warp-uniform property of the work item): no clamp logic, no divisions.
template <typename R, int O, bool GRAD>
__device__ __forceinline__ int axis_weights_interior(const AxisParams<R>& a, R x, int k, R* w, R* dw) {
    constexpr int WIN = 2 * O > 0 ? 2 * O : 1;
    R fs;
    const int span = locate_uniform_interior<R, O>(a, x, fs);
    if ((GRAD || k == 0) && a.unit_ok) {
        // the span is exact (compared against the reference's knot values); the weights only need
        // the offset inside the cell (well-conditioned ranges only, see Grid::params)
        const R u = (x - uniform_knot<R>(a, fs)) * a.inv_dx;
        basis_unit<R, O, GRAD>(u, a.inv_dx, w, dw);
    } else {
        R tk[WIN];
        uniform_knot_window<R, O>(a, fs, tk);
        if (GRAD || k == 0) basis_uniform<R, O, GRAD>(tk, x, a.inv_dx, w, dw);
        else deriv_weights<R, O>(tk, x, k, w);
    }
    return span - O;
}

// A tile is "interior" on an axis when the axis is uniform and all cells c0 of the tile
// satisfy O <= c0 <= K-3O-2 (periodic axes: always), see bspl_device.cuh.
template <typename R, int O>
__device__ __forceinline__ bool tile_interior(const AxisParams<R>& a, int t) {
    constexpr int T = tile_edge<O>();
    if (a.t != nullptr) return false;
    if (a.periodic) return true;
    return t * T >= O && t * T + T - 1 <= a.K - 3 * O - 2;
}

template <typename R, int O, bool GRAD>
__global__ void __launch_bounds__(kEvalThreads, 2)
    eval_binned_kernel(const BinEvalParams<R> p, const __grid_constant__ CUtensorMap tmap) {
    constexpr int T = tile_edge<O>();
    constexpr int BE = brick_edge<O>();
    constexpr int BP = brick_pitch<R, O>();
    constexpr int W = O + 1;
    constexpr int NOUT = GRAD ? 4 : 1;
    constexpr bool DUAL = dual_brick<O>();
    constexpr uint32_t kBrickBytes = BE * BE * BP * sizeof(R);
    // the shifted copy starts half the banks further, so that z offsets 2k and 2k+1 (same chunk index, one
    // in each copy) do not collide
    constexpr uint32_t kBrickStride = (kBrickBytes + 127) / 128 * 128 + 64;
    using R2 = typename std::conditional<sizeof(R) == 8, double2, float2>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    R* brick = reinterpret_cast<R*>(smem_raw);
    R* brick_odd = reinterpret_cast<R*>(smem_raw + kBrickStride);  // brick_odd[z] == brick[z + 1]
    __shared__ uint64_t bar;
    __shared__ uint32_t s_item;
    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    uint32_t parity = 0;
    const uint32_t n_work = *p.n_work;

    for (;;) {
        if (tid == 0) s_item = atomicAdd(p.next_item, 1u);
        __syncthreads();  // publishes s_item; every thread is past its reads of the previous brick
        const uint32_t item = s_item;
        __syncthreads();  // s_item is rewritten by thread 0 at the top of the next iteration
        if (item >= n_work) break;
        const uint32_t tile = p.work[3 * item], begin = p.work[3 * item + 1], end = p.work[3 * item + 2];
        const int tz = tile % p.ntile[2], ty = (tile / p.ntile[2]) % p.ntile[1], tx = tile / (p.ntile[2] * p.ntile[1]);
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(&bar, kBrickBytes);
            tma_load_3d(brick, &tmap, &bar, tz * T, ty * T, tx * T);
        }
        const bool in0 = tile_interior<R, O>(p.ax[0], tx), in1 = tile_interior<R, O>(p.ax[1], ty),
                   in2 = tile_interior<R, O>(p.ax[2], tz);
        // fetch this thread's first query while the brick is in flight
        uint32_t i = begin + tid;
        Rec<R> cur;
        cur.x = cur.y = cur.z = R(0); cur.idx = 0;
        if (i < end) cur = p.rec[i];
        mbar_wait(&bar, parity);
        parity ^= 1;
        if (DUAL) {
            // second copy shifted by one element (TMA box origins must stay 16-byte aligned,
            // so the shift is done here, shared -> shared)
            for (int e = tid; e < BE * BE * BP - 1; e += kEvalThreads) brick_odd[e] = brick[e + 1];
            __syncthreads();
        }
        while (i < end) {
            // software pipeline: the next record is in flight while this one is evaluated
            const uint32_t i_n = i + kEvalThreads;
            Rec<R> nxt = cur;
            if (i_n < end) nxt = p.rec[i_n];
            const R x0 = cur.x, x1 = cur.y, x2 = cur.z;
            const long long idx = static_cast<long long>(cur.idx);
            R w[3][W], dw[GRAD ? 3 : 1][W];
            int cx, cy, cz;
            if (in0) cx = axis_weights_interior<R, O, GRAD>(p.ax[0], x0, GRAD ? 0 : p.deriv[0], w[0], dw[0]);
            else cx = axis_weights<R, O, GRAD>(p.ax[0], x0, GRAD ? 0 : p.deriv[0], w[0], dw[0]);
            if (in1) cy = axis_weights_interior<R, O, GRAD>(p.ax[1], x1, GRAD ? 0 : p.deriv[1], w[1], dw[GRAD ? 1 : 0]);
            else cy = axis_weights<R, O, GRAD>(p.ax[1], x1, GRAD ? 0 : p.deriv[1], w[1], dw[GRAD ? 1 : 0]);
            if (in2) cz = axis_weights_interior<R, O, GRAD>(p.ax[2], x2, GRAD ? 0 : p.deriv[2], w[2], dw[GRAD ? 2 : 0]);
            else cz = axis_weights<R, O, GRAD>(p.ax[2], x2, GRAD ? 0 : p.deriv[2], w[2], dw[GRAD ? 2 : 0]);
            cx -= tx * T; cy -= ty * T; cz -= tz * T;
            const R* c = (DUAL && (cz & 1)) ? brick_odd + (cx * BE + cy) * BP + (cz - 1)
                                            : brick + (cx * BE + cy) * BP + cz;
            R v = R(0), g0 = R(0), g1 = R(0), g2 = R(0);
#pragma unroll
            for (int a = 0; a < W; ++a) {
                R bi = R(0), bi1 = R(0), bi2 = R(0);
#pragma unroll
                for (int b = 0; b < W; ++b) {
                    const R* row = c + (a * BE + b) * BP;
                    R s = R(0), s2 = R(0);
                    if (DUAL) {
#pragma unroll
                        for (int k = 0; k < W / 2; ++k) {
                            const R2 cv = reinterpret_cast<const R2*>(row)[k];
                            s += cv.x * w[2][2 * k];
                            s += cv.y * w[2][2 * k + 1];
                            if (GRAD) {
                                s2 += cv.x * dw[GRAD ? 2 : 0][2 * k];
                                s2 += cv.y * dw[GRAD ? 2 : 0][2 * k + 1];
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < W; ++k) {
                            const R cv = row[k];
                            s += cv * w[2][k];
                            if (GRAD) s2 += cv * dw[GRAD ? 2 : 0][k];
                        }
                    }
                    bi += s * w[1][b];
                    if (GRAD) {
                        bi1 += s * dw[GRAD ? 1 : 0][b];
                        bi2 += s2 * w[1][b];
                    }
                }
                v += bi * w[0][a];
                if (GRAD) {
                    g0 += bi * dw[0][a];
                    g1 += bi1 * w[0][a];
                    g2 += bi2 * w[0][a];
                }
            }
            R* o = p.out + idx * NOUT;
            if (GRAD) {
                // value and gradient leave as one aligned 4-element store (a whole 32-byte sector for fp64)
                Out4<R> r4;
                r4.v = v; r4.g0 = g0; r4.g1 = g1; r4.g2 = g2;
                *reinterpret_cast<Out4<R>*>(o) = r4;
            } else {
                o[0] = v;
            }
            i = i_n; cur = nxt;
        }
    }
}

// ---- host side ----------------------------------------------------------------

template <typename R, int O>
cudaError_t eval_binned_O(const EvalArgs<R>& a, const BinnedScratch& sc, int phases, cudaStream_t s) {
    constexpr int T = tile_edge<O>();
    constexpr int BE = brick_edge<O>();
    constexpr int BP = brick_pitch<R, O>();
    BinParams<R> bp;
    BinEvalParams<R> ep;
    int n_tiles = 1;
    for (int d = 0; d < 3; ++d) {
        bp.ax[d] = a.ax[d];
        ep.ax[d] = a.ax[d];
        const int cells = a.ax[d].K - 2 * O - 1;  // distinct values of span - O
        bp.ntile[d] = ep.ntile[d] = (cells + T - 1) / T;
        n_tiles *= bp.ntile[d];
        ep.deriv[d] = a.deriv[d];
    }
    if (n_tiles > sc.max_tiles) return cudaErrorInvalidValue;
    constexpr int kBinsPerTile = T * T;
    const long long n_bins = static_cast<long long>(n_tiles) * kBinsPerTile;
    bp.pts = a.pts; bp.q = a.q; bp.n_tiles = n_tiles;

    cudaError_t e = cudaSuccess;
    if (phases & kBinnedSort) {
        e = cudaMemsetAsync(sc.counts, 0, sizeof(uint32_t) * n_bins, s);
        if (e != cudaSuccess) return e;
        const int grid = kSMs * 4;
        key_count_kernel<R, O><<<grid, 512, 0, s>>>(bp, sc.counts);
        const int tgrid = (n_tiles * 32 + 255) / 256;
        tile_totals_kernel<<<tgrid, 256, 0, s>>>(sc.counts, n_tiles, kBinsPerTile, sc.tile_total);
        plan_kernel<<<1, 1024, 0, s>>>(sc.tile_total, n_tiles, sc.tile_off, sc.work, sc.n_work, sc.next_item);
        key_cursor_kernel<<<tgrid, 256, 0, s>>>(sc.counts, n_tiles, kBinsPerTile, sc.tile_off, sc.cursor);
        scatter_kernel<R, O><<<grid, 512, 0, s>>>(bp, sc.cursor, static_cast<Rec<R>*>(sc.rec));
        count_launch(5);
        e = cudaGetLastError();
        if (e != cudaSuccess || !(phases & kBinnedEval)) return e;
    }
    // tensor map over the padded coefficient array: dims (z, y, x), row-major
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return cudaErrorNotSupported;
    CUtensorMap tmap;
    const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(a.ax[1].stride),  // padded z extent == y stride
                                static_cast<cuuint64_t>(a.ax[0].stride / a.ax[1].stride),
                                static_cast<cuuint64_t>(a.field_stride / a.ax[0].stride)};
    const cuuint64_t gstr[2] = {static_cast<cuuint64_t>(a.ax[1].stride) * sizeof(R),
                                static_cast<cuuint64_t>(a.ax[0].stride) * sizeof(R)};
    const cuuint32_t box[3] = {BP, BE, BE};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult cr = enc(&tmap, sizeof(R) == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                            3, const_cast<R*>(a.coef), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;

    // the work counter is consumed by every evaluation of a (possibly reused) sorted batch
    e = cudaMemsetAsync(sc.next_item, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;

    ep.rec = static_cast<const Rec<R>*>(sc.rec); ep.out = a.out; ep.work = sc.work; ep.n_work = sc.n_work;
    ep.next_item = sc.next_item;
    constexpr int kOne = (BE * BE * BP * int(sizeof(R)) + 127) / 128 * 128;
    constexpr int smem = dual_brick<O>() ? 2 * kOne + 64 : kOne;
    if (a.mode == kValueGrad) {
        auto k = eval_binned_kernel<R, O, true>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<kSMs * 2, kEvalThreads, smem, s>>>(ep, tmap);
    } else {
        auto k = eval_binned_kernel<R, O, false>;
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        k<<<kSMs * 2, kEvalThreads, smem, s>>>(ep, tmap);
    }
    count_launch(1);
    return cudaGetLastError();
}

}  // namespace

size_t binned_scratch_bytes(long long q, int max_tiles, size_t* offsets) {
    // layout: key_of[q] | rec[q] (32 bytes each) | counts | cursor | work[3*(q/kChunk + max_tiles)] | n_work, next | ...
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    offsets[0] = take(256);  // (formerly the per-query key array)
    offsets[1] = take(32 * static_cast<size_t>(q));
    offsets[2] = take(sizeof(uint32_t) * max_tiles * 256);  // one counter per (tile, x, y) key
    offsets[3] = take(sizeof(uint32_t) * max_tiles * 256);
    offsets[4] = take(sizeof(uint32_t) * 3 * (q / kChunk + max_tiles + 1));
    offsets[5] = take(256);
    offsets[6] = take(sizeof(uint32_t) * max_tiles);
    offsets[7] = take(sizeof(uint32_t) * max_tiles);
    return off;
}

BinnedScratch binned_scratch_view(void* base, long long q, int max_tiles) {
    size_t off[8];
    binned_scratch_bytes(q, max_tiles, off);
    unsigned char* b = static_cast<unsigned char*>(base);
    BinnedScratch sc;
    sc.tile_of = reinterpret_cast<uint32_t*>(b + off[0]);
    sc.rec = b + off[1];
    sc.counts = reinterpret_cast<uint32_t*>(b + off[2]);
    sc.cursor = reinterpret_cast<uint32_t*>(b + off[3]);
    sc.work = reinterpret_cast<uint32_t*>(b + off[4]);
    sc.n_work = reinterpret_cast<uint32_t*>(b + off[5]);
    sc.next_item = sc.n_work + 1;
    sc.tile_total = reinterpret_cast<uint32_t*>(b + off[6]);
    sc.tile_off = reinterpret_cast<uint32_t*>(b + off[7]);
    sc.max_tiles = max_tiles;
    return sc;
}

template <typename R>
int binned_tile_count(const EvalArgs<R>& a) {
    if (a.dim != 3 || a.order > 5) return 0;
    const int T = tile_edge_rt(a.order);
    long long n = 1;
    for (int d = 0; d < 3; ++d) n *= (a.ax[d].K - 2 * a.order - 1 + T - 1) / T;
    return n > (1 << 22) ? 0 : static_cast<int>(n);
}

template <typename R>
cudaError_t launch_eval_binned(const EvalArgs<R>& a, const BinnedScratch& sc, cudaStream_t s, int phases) {
    if (a.q <= 0) return cudaSuccess;
    if (a.dim != 3 || a.n_fields != 1 || a.q >= (1ll << 32)) return cudaErrorInvalidValue;
    switch (a.order) {
        case 0: return eval_binned_O<R, 0>(a, sc, phases, s);
        case 1: return eval_binned_O<R, 1>(a, sc, phases, s);
        case 2: return eval_binned_O<R, 2>(a, sc, phases, s);
        case 3: return eval_binned_O<R, 3>(a, sc, phases, s);
        case 4: return eval_binned_O<R, 4>(a, sc, phases, s);
        case 5: return eval_binned_O<R, 5>(a, sc, phases, s);
        default: return cudaErrorInvalidValue;
    }
}

template cudaError_t launch_eval_binned<double>(const EvalArgs<double>&, const BinnedScratch&, cudaStream_t, int);
template cudaError_t launch_eval_binned<float>(const EvalArgs<float>&, const BinnedScratch&, cudaStream_t, int);
template int binned_tile_count<double>(const EvalArgs<double>&);
template int binned_tile_count<float>(const EvalArgs<float>&);

}  // namespace bspl
