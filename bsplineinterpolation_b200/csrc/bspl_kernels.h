// Host-visible launch interface between the C-ABI layer (bspl_capi.cu) and the
// kernel translation units.  Internal; the public surface is include/bspline_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bspl_device.cuh"

namespace bspl {

enum EvalMode { kValue = 0, kValueGrad = 1 };

template <typename R>
struct EvalArgs {
    int dim, order;
    AxisParams<R> ax[kMaxDim];
    const R* coef;            // padded coefficient array of the first field to evaluate
    long long field_stride;   // elements between consecutive fields
    int n_fields;             // fields evaluated by this launch (out is [n_fields][q][n_out])
    const R* pts;             // [q][dim]
    R* out;
    long long q;
    int deriv[kMaxDim];       // kValue only
    int mode;
};

// Direct gather: one query per thread straight from the padded global array.
template <typename R>
cudaError_t launch_eval_direct(const EvalArgs<R>& a, cudaStream_t s);

// Cell-binned path (3-D): tile the queries, stage each tile's coefficient brick in
// shared memory with TMA, evaluate out of shared memory (bspl_binned.cu).
struct BinnedScratch {
    uint32_t* tile_of;
    void* rec;  // sorted query records (coordinates + original index)
    uint32_t* counts;
    uint32_t* cursor;
    uint32_t* work;
    uint32_t* n_work;
    uint32_t* next_item;
    uint32_t* tile_total;
    uint32_t* tile_off;
    int max_tiles;
};
size_t binned_scratch_bytes(long long q, int max_tiles, size_t* offsets /*[8]*/);
BinnedScratch binned_scratch_view(void* base, long long q, int max_tiles);
template <typename R>
int binned_tile_count(const EvalArgs<R>& a);  // 0 when the binned path does not apply
// phases: sort the queries into `sc` (needs a.pts), evaluate the sorted batch (needs a.out), or both
enum { kBinnedSort = 1, kBinnedEval = 2 };
template <typename R>
cudaError_t launch_eval_binned(const EvalArgs<R>& a, const BinnedScratch& sc, cudaStream_t s,
                               int phases = kBinnedSort | kBinnedEval);

// Many fields of one small 2-D mesh at one query set: fields streamed through shared memory
// with bulk asynchronous copies, weights kept in registers (bspl_fields.cu).
template <typename R>
bool fields_smem_eligible(const EvalArgs<R>& a);
template <typename R>
cudaError_t launch_eval_fields_smem(const EvalArgs<R>& a, cudaStream_t s);

// Many fields at one query set as a per-cell contraction with query-major results (bspl_contract.cu).
struct FieldsScratch {
    uint32_t* counts;      // [n_keys] queries per cell
    uint32_t* cursor;      // [n_keys]
    uint32_t* idx_sorted;  // [q] original index of the query at each sorted position
    void* w_sorted;        // [q][K] tensor-product weights in cell order
    uint32_t* work;        // (key, begin, end) triples
    uint32_t* n_work;
    uint32_t* next_unit;
};
template <typename R>
struct FieldsContractArgs {
    int dim, order, K;          // K = (order+1)^dim stencil terms
    AxisParams<R> ax[kMaxDim];
    int deriv[kMaxDim];
    int n_keys;                 // padded elements per field: a cell's key is the offset of its first control point
    int chunk;                  // queries per work item (fields_query_chunk)
    const R* pts;               // [q][dim]                         (sort phase)
    long long q;
    const R* coef_t;            // [n_keys][n_fields] field-minor   (contract phase)
    int n_fields;
    int field_begin, field_end; // fields evaluated by this launch
    R* out;                     // (query, field) at query * out_stride + field - field_begin
    long long out_stride;
    int off[64];                // offset of stencil term k from the cell's first control point, last axis fastest
};
int fields_query_chunk(int K);
size_t fields_scratch_bytes(long long q, int n_keys, int K, size_t elem, size_t* offsets /*[6]*/);
FieldsScratch fields_scratch_view(void* base, long long q, int n_keys, int K, size_t elem);
bool fields_contract_supported(int dim, int order);
template <typename R>
cudaError_t launch_fields_sort(const FieldsContractArgs<R>& a, const FieldsScratch& sc, cudaStream_t s);
// cudaErrorNotSupported: alignment of out / coef_t / n_fields does not allow the vector accesses
template <typename R>
cudaError_t launch_fields_contract(const FieldsContractArgs<R>& a, const FieldsScratch& sc, cudaStream_t s);

// span - order per axis, int32 [q][dim]
template <typename R>
cudaError_t launch_locate(const EvalArgs<R>& a, int32_t* cell, cudaStream_t s);

// ---- control-point solve ----------------------------------------------------

// Device-resident LU factors of one axis in row form (see bspl_solve.cu).
template <typename R>
struct AxisLU {
    int n, p, q, cyclic;
    const R* L;       // [n][p]  L(i, i-p+m)
    const R* U;       // [n][q]  U(i, i+1+m)
    const R* diag;    // [n]     U(i, i)
    const R* rdiag;   // [n]     1 / U(i, i) refined exactly as the fast path of CUDA's IEEE division refines it
                      //         (fill_refined_reciprocals); fp64 only, nullptr otherwise
    const R* bottom;  // [n][q]  side(n-q+r, j) at [j*q + r]          (cyclic)
    const R* right;   // [n][p]  side(i, n-p+c) at [i*p + c]          (cyclic)
    // Row-packed copies for the tiled sweeps, which fetch the factor rows of a whole tile with one bulk copy
    // (launch_pack_factors; n + 32 rows, the padding rows are identity rows):
    //   fwd_pack[j] = {L(j, .)[p], bottom(., j)[p] if cyclic}          bwd_pack[j] = {U(j, .)[p], right(j, .)[p] if cyclic,
    //                                                                                  U(j, j), rdiag[j]}
    // nullptr for compact storage (head < n).
    const R* fwd_pack;
    const R* bwd_pack;
    int bottom_len;   // entries j >= bottom_len are exactly zero
    int right_len;    // entries i >= right_len are exactly zero (rows above the main band only)
    int bottom_sig;   // entries j >= bottom_sig are below 1e-30 of the largest (chunked sweeps only)
    // Compact storage of long uniform axes (bspl_host.h: build_axis_factor): L, U and diag hold
    // rows [0, head) and the last rows of the matrix; the `skip` rows cut out in between all
    // equal row head-1.  Stored in full: head == n, skip == 0.
    int head, skip;
    __host__ __device__ __forceinline__ long long row(int j) const {
        return j < head ? j : (j >= head + skip ? j - skip : head - 1);
    }
};

// Geometry of one sweep over a (field, axis0, axis1, axis2) array: lines run
// along `axis`, the remaining (up to three) dimensions enumerate the lines.
struct SweepGeom {
    int n;                   // line length
    long long line_stride;   // element stride along the line
    int m[3];                // sizes of the other dimensions, m[2] fastest
    long long ms[3];         // their element strides
};

// chunk == 0: exact sequential sweep, one thread per line (bit-identical to the reference).
// chunk  > 0: chunk-parallel sweep (few long lines); scratch holds y (scratch_y_elems) followed
// by lines*max(P,1) tail values.
struct SweepPlan {
    int chunk;
    int window;
    void* scratch;
    long long scratch_y_elems;
};
SweepPlan plan_sweep(int n, long long lines, int window, int cyclic, int bottom_sig);
template <typename R>
cudaError_t launch_sweep(const AxisLU<R>& lu, const SweepGeom& g, R* data, const SweepPlan& plan, cudaStream_t s);
// out[i] = reciprocal of diag[i] for AxisLU::rdiag (device pointers)
cudaError_t fill_refined_reciprocals(const double* diag, double* out, long long n, cudaStream_t s);
// AxisLU::fwd_pack / bwd_pack from the row tables of `lu` (device pointers; rows = n + padding)
template <typename R>
cudaError_t launch_pack_factors(const AxisLU<R>& lu, long long rows, R* fwd_pack, R* bwd_pack, cudaStream_t s);
// Collocation rows + band LU of a long non-uniform axis on the device (bspl_factor.cu).  assemble: coords, knots
// device arrays -> band [n][2 bw + 1] (bw = order-1, or order/2 on a periodic axis, whose wrapping entries are left
// out).  factor: band -> L, U [n][max(bw,1)], diag [n] in AxisLU's row form for rows >= first_row (warm-ups never
// reach below it); check [device_band_factor_check_elems] is scratch; *flag != 0 afterwards: the chunks did not
// agree bit for bit at their seams -- discard the result.
template <typename R>
cudaError_t launch_device_band_assemble(int order, int periodic, long long n, long long K, const R* coords, const R* knots,
                                        R* band, cudaStream_t s);
template <typename R>
cudaError_t launch_device_band_factor(int bw, long long n, long long first_row, const R* band, R* check, R* L, R* U,
                                      R* diag, int* flag, int chunk, int window, cudaStream_t s);
size_t device_band_factor_check_elems(long long n, int bw, int chunk);
inline int fwd_pack_width(int p, int cyclic) { return (cyclic ? 2 * p : p) > 0 ? (cyclic ? 2 * p : p) : 1; }
inline int bwd_pack_width(int p, int cyclic) { return (cyclic ? 2 * p : p) + 2; }

// First sweep of the separable solve fused with the copy out of the caller's mesh: lines along
// the contiguous axis are read from `src` (line space strides src_ms, same extents g.m) and the
// solved lines land in `dst` (geometry g), shifted cyclically by shift[k] along g.m[k] and, when
// `rotate` is set (periodic line axis), by the band width along the line itself
// (InterpolationTemplate.hpp:451-462).  cudaErrorNotSupported: geometry not addressable by TMA
// (odd strides, unaligned base) -- the caller takes another route.
template <typename R>
cudaError_t launch_sweep_contig_from(const AxisLU<R>& lu, const SweepGeom& g, const R* src, const long long* src_ms,
                                     const int* shift, int rotate, R* dst, cudaStream_t s);

// f (compact, [fields][n0][n1][n2]) -> padded coefficient array, rotating each
// periodic axis by +shift[d] (InterpolationTemplate.hpp:451-462).
struct CopyGeom {
    int dim;
    int n[kMaxDim];
    int shift[kMaxDim];
    long long dst_stride[kMaxDim];
    long long src_field_stride, dst_field_stride;
    long long fields;
};
template <typename R>
cudaError_t launch_rotate_copy(const CopyGeom& g, const R* src, R* dst, cudaStream_t s);
// fill the `ghost[d]` wrap-around cells of every periodic axis (cells n..n+ghost-1 := 0..ghost-1)
struct GhostGeom {
    int dim;
    int n[kMaxDim];
    int ghost[kMaxDim];
    long long stride[kMaxDim];
    long long field_stride;
    long long fields;
};
template <typename R>
cudaError_t launch_fill_ghosts(const GhostGeom& g, R* data, cudaStream_t s);
// padded -> compact copy of one field (for control_points())
template <typename R>
cudaError_t launch_unpad_copy(const CopyGeom& g, const R* src_padded, R* dst_compact, cudaStream_t s);

// Sweep fused with the re-shard of the slab-sharded multi-GPU solve: the forward pass runs in
// place on the local slab, the backward pass stores every solved row straight into the buffer of
// the rank that owns it after the exchange (peer-mapped device memory, NVLink stores) -- no pack,
// no separate all-to-all, no unpack.
constexpr int kMaxPeers = 8;
template <typename R>
struct ExchangeDest {
    int n_ranks;
    int split[kMaxPeers + 1];   // rows [split[r], split[r+1]) of every line belong to rank r
    R* base[kMaxPeers];         // rank r's buffer, already offset to this rank's block in it
    long long ms[kMaxPeers][3]; // strides of the three outer line indices in rank r's buffer
    long long ls[kMaxPeers];    // stride between consecutive rows in rank r's buffer
    // the middle outer index lands at (i1 + i1_offset) mod i1_mod in the destination (i1_mod == 0: as it is):
    // the slab offset of this rank plus the rotation of a periodic slab axis (InterpolationTemplate.hpp:455-459)
    int i1_offset, i1_mod;
};
template <typename R>
cudaError_t launch_sweep_exchange(const AxisLU<R>& lu, const SweepGeom& g, R* data, const ExchangeDest<R>& dest,
                                  cudaStream_t s);

// Stream-ordered barrier between the ranks of one node (one process per GPU): every rank's kernel
// bumps its own epoch counter, stores the epoch into its slot of every peer's flag array
// (peer-mapped memory, release at system scope) and waits until all of its own slots have
// reached the epoch.  No host synchronisation; gives up after `timeout_cycles` and sets *status.
struct RankBarrier {
    int rank, n_ranks;
    unsigned int* flags[kMaxPeers];  // flags[r]: rank r's array of kMaxPeers slots (this rank's own for r == rank)
    unsigned int* epoch;             // this rank's barrier count (device memory)
    int* status;                     // set to 1 on timeout (device memory)
    long long timeout_cycles;
};
cudaError_t launch_rank_barrier(const RankBarrier& b, cudaStream_t s);

// Batched transpose of the last two axes through shared memory:
//   dst[(b0, b1), rot_q(q), rot_p(p)] = src[(b0, b1), p, q]     (src q-contiguous, dst p-contiguous)
// with optional rotation of every index by +shift (mod extent) on the destination side.
// (b0, b1) are two outer batch indices with independent strides (fields, leading axis).
struct TransposeGeom {
    int nb0, nb1, np, nq;
    long long src_b0, src_b1, src_p;            // src element (b0,b1,p,q) at b0*src_b0 + b1*src_b1 + p*src_p + q
    long long dst_b0, dst_b1, dst_q;            // dst element at b0*dst_b0 + b1'*dst_b1 + q'*dst_q + p'
    int shift_b1, shift_p, shift_q;
};
template <typename R>
cudaError_t launch_transpose(const TransposeGeom& g, const R* src, R* dst, cudaStream_t s);

// 0: automatic, 1: thread-per-line sweeps only, 2: the L2-resident tiled sweep whenever TMA can address the lines
void set_sweep_path(int path);

void count_launch(int n = 1);

}  // namespace bspl
