/* bspline_b200 -- C ABI of the B200 (sm_100a) implementation of the
 * BSplineInterpolation hot path: batched evaluation of
 * intp::InterpolationFunction<T, D, Order, U> and the separable control-point
 * solve that produces its coefficients.
 *
 * The reference (12ff54e/BSplineInterpolation, header-only C++) has no FFI;
 * its boundary is the template API in src/include/Interpolation.hpp and
 * src/include/InterpolationTemplate.hpp.  Each entry point below names the
 * reference interface it stands in for (file:line in the reference checkout).
 * <T, D, Order, U> are compile-time there and run-time here (dtype, dim,
 * order); T == U (real scalars).
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns a bspl_status and
 *     records a message retrievable with bspl_last_error() (thread-local);
 *   - meshes and control points are row-major, last index fastest
 *     (Mesh.hpp:241-246); query points are [q][dim] (array of DimArray<coord>,
 *     Interpolation.hpp:144);
 *   - `on_device` != 0: pts/out/f are device pointers on the handle's device
 *     and the work is enqueued on `stream` (a cudaStream_t, NULL = default
 *     stream) without synchronising; == 0: host pointers, the call copies in,
 *     runs, copies out and returns when the result is in `out`;
 *   - there is NO CPU fallback: without a CUDA device every compute call
 *     returns BSPL_ERR_CUDA.
 */
#ifndef BSPLINE_B200_H
#define BSPLINE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSPL_MAX_DIM 4
#define BSPL_MAX_ORDER 7

typedef enum {
    BSPL_OK = 0,
    BSPL_ERR_INVALID = 1,     /* bad argument (INTP_ASSERT -> std::runtime_error, util.hpp:247-258) */
    BSPL_ERR_CUDA = 2,        /* CUDA runtime failure / no device */
    BSPL_ERR_DOMAIN = 3,      /* coordinate out of range (std::domain_error, Interpolation.hpp:482-491) */
    BSPL_ERR_ALLOC = 4,       /* std::bad_alloc */
    BSPL_ERR_UNSUPPORTED = 5  /* dim/order outside the instantiated set */
} bspl_status;

typedef enum { BSPL_F64 = 0, BSPL_F32 = 1 } bspl_dtype;

/* InterpolationFunctionTemplate<T,D,O,U> (InterpolationTemplate.hpp:32-581):
 * per-axis knot vectors and the LU factors of the collocation matrices, on
 * the device, reusable for any number of fields on the same mesh. */
typedef struct bspl_template bspl_template;

/* InterpolationFunction<T,D,O,U> (Interpolation.hpp:17-507): knots plus the
 * control points of n_fields >= 1 fields, device resident. */
typedef struct bspl_function bspl_function;

/* InterpolationFunctionTemplate ctor (InterpolationTemplate.hpp:60-79) ->
 * create_knot_vector_ (Interpolation.hpp:322-362 uniform, :365-464
 * non-uniform) + build_solver_ (InterpolationTemplate.hpp:254-446, BandLU
 * factorisation BandLU.hpp:103-118 / :159-213).  n[d] = data points on axis d
 * (INTP_PERIODIC_NO_DUMMY_POINT semantics: the closing sample of a periodic
 * axis is implicit).  coords == NULL or coords[d] == NULL: uniform axis on
 * [lo[d], hi[d]]; otherwise coords[d] holds n[d] (+1 if periodic) increasing
 * abscissae.  device = CUDA ordinal. */
int bspl_template_create(bspl_dtype dtype, int dim, int order, const int64_t* n,
                         const int* periodic, const double* lo, const double* hi,
                         const double* const* coords, int device, bspl_template** out);
void bspl_template_destroy(bspl_template* t);
/* Solver of one axis (solvers_[axis], InterpolationTemplate.hpp:515): half bandwidth p == q
 * (:291), whether it is the cyclic kind (ExtendedBandMatrix), and whether its collocation matrix
 * and LU factors were built on the device (long non-uniform, non-periodic axes; everything else
 * is built on the host, long uniform axes in O(1)).  Any out pointer may be NULL. */
int bspl_template_axis_info(const bspl_template* t, int axis, int* band, int* cyclic, int* built_on_device);

/* interpolate(mesh) const& (InterpolationTemplate.hpp:118-125) ->
 * solve_for_control_points_ (:448-580) for n_fields meshes stored back to back
 * ([n_fields][n0]...[nD-1]); returns a new function. */
int bspl_template_interpolate(const bspl_template* t, const void* f, int64_t n_fields,
                              int on_device, void* stream, bspl_function** out);
/* interpolate(function_type&, mesh) (InterpolationTemplate.hpp:136-143): reuse
 * the storage of an existing function of the same template. */
int bspl_template_interpolate_into(const bspl_template* t, bspl_function* fn, const void* f,
                                   int64_t n_fields, int on_device, void* stream);

/* One stage of solve_for_control_points_ on caller-owned DEVICE memory: the banded /
 * cyclic solve of template axis `axis` (solvers_[axis], InterpolationTemplate.hpp:515)
 * applied in place to every line of a strided array.  A line has n[axis] elements
 * `line_stride` apart; lines are enumerated by up to three outer indices i0 < m[0],
 * i1 < m[1], i2 < m[2] at element offsets i0*ms[0] + i1*ms[1] + i2*ms[2] (m[2] should be
 * the contiguous-most).  The periodic-axis rotation of the right-hand side
 * (:451-462) is NOT applied here.  Building block of the slab-sharded multi-GPU solve
 * (sweep local axes, all-to-all, sweep the remaining axis). */
int bspl_template_sweep_axis(const bspl_template* t, int axis, void* data, const int64_t* m,
                             const int64_t* ms, int64_t line_stride, void* stream);
/* bspl_template_sweep_axis fused with the re-shard of the slab-sharded solve: the forward
 * substitution runs in place on `data`; the backward substitution stores row j of every line
 * into the buffer of the rank that owns it, rank r owning rows [split[r], split[r+1]).  The
 * element (i0, i1, i2, j) lands at peer_base[r] + i0*peer_ms[3r] + i1*peer_ms[3r+1] +
 * i2*peer_ms[3r+2] + (j - split[r]) * peer_ls[r].  peer_base[r] is device memory of GPU
 * peer_device[r] mapped into this process (CUDA IPC); peer access is enabled on demand.
 * `data` holds the forward-substituted values afterwards and is scratch.  The caller
 * synchronises the ranks before reading its own buffer.  n_ranks <= 8. */
int bspl_template_sweep_axis_exchange(const bspl_template* t, int axis, void* data, const int64_t* m,
                                      const int64_t* ms, int64_t line_stride, int n_ranks,
                                      const int64_t* split, void* const* peer_base,
                                      const int* peer_device, const int64_t* peer_ms,
                                      const int64_t* peer_ls, void* stream);
/* ---- slab-sharded solve of ONE 3-D field over the GPUs of a node ---------------------------------
 * No reference counterpart (the reference is one process; its parallel unit is the
 * INTP_MULTITHREAD line loop, InterpolationTemplate.hpp:547-572).  The arithmetic per line is
 * solve_for_control_points_'s (:448-580); its per-axis solves commute, so a mesh cut into slabs of
 * axis 0 sweeps axes 2 and 1 locally, re-shards to slabs of axis 1 and sweeps axis 0 there.  One
 * process per GPU; rank r of n_ranks owns planes [slab0[r], slab0[r+1]) of axis 0 before and
 * [slab1[r], slab1[r+1]) of axis 1 after the exchange (even split, remainder to the low ranks).
 * Every rank creates its plan from its own template (same mesh, its own device), publishes the
 * 64-byte handle of its receive buffer to the others by any means, and connects. */
typedef struct bspl_sharded_solve bspl_sharded_solve;
int bspl_sharded_solve_create(const bspl_template* t, int rank, int n_ranks, bspl_sharded_solve** out);
void bspl_sharded_solve_destroy(bspl_sharded_solve* s);   /* the template must outlive the plan */
/* slab boundaries, n_ranks + 1 entries each (either may be NULL) */
int bspl_sharded_solve_layout(const bspl_sharded_solve* s, int64_t* slab0, int64_t* slab1);
int bspl_sharded_solve_handle(const bspl_sharded_solve* s, unsigned char handle_out[64]);
/* handles: n_ranks * 64 bytes, entry r from rank r (the own entry is ignored; NULL when n_ranks == 1) */
int bspl_sharded_solve_connect(bspl_sharded_solve* s, const unsigned char* handles);
/* The whole solve on `stream`, no host synchronisation: f_slab is this rank's DEVICE slab
 * [slab0 planes][n1][n2] of the mesh; *ctrl_slab receives a device pointer to this rank's slab
 * [n0][slab1 planes][n2] of the control points, owned by the plan and valid until its next run.
 * The axis-1 sweep stores its solved rows straight into the owners' buffers (peer-mapped memory,
 * NVLink stores) -- sweep and exchange are one kernel -- and the ranks meet at two stream-ordered
 * flag barriers.  Collective: every rank calls it the same number of times. */
int bspl_sharded_solve_run(bspl_sharded_solve* s, const void* f_slab, void** ctrl_slab, void* stream);
/* The same solve around a caller-run all-to-all (NCCL): pack() sweeps axes 2 and 1 and leaves the
 * blocks for the ranks back to back in *send (send_counts[r] elements for rank r); the caller
 * exchanges them into *recv (recv_counts[r] elements from rank r, in rank order) on the same
 * stream; finish() sweeps axis 0 and returns the slab as run() does. */
int bspl_sharded_solve_pack(bspl_sharded_solve* s, const void* f_slab, void** send, void** recv,
                            int64_t* send_counts, int64_t* recv_counts, void* stream);
int bspl_sharded_solve_finish(bspl_sharded_solve* s, void** ctrl_slab, void* stream);
/* *timed_out != 0: a flag barrier gave up waiting for a peer (~10 s); results are invalid.  Synchronises. */
int bspl_sharded_solve_status(bspl_sharded_solve* s, int* timed_out);

/* Device buffers shareable between the ranks of one node (CUDA IPC), for the exchange above:
 * alloc returns device memory of `device` and a 64-byte handle to send to the peers; open maps
 * a peer's buffer for access from kernels running on `device` (the ACCESSING device).  Close
 * every opened mapping before the owner frees the buffer. */
int bspl_ipc_alloc(int device, int64_t bytes, void** dptr, unsigned char handle_out[64]);
int bspl_ipc_open(int device, const unsigned char handle[64], void** dptr);
int bspl_ipc_close(int device, void* dptr);
int bspl_ipc_free(int device, void* dptr);
/* Wrap already-solved plain control points ([n_fields][n0]...[nD-1], host or device) into
 * a function of this template (load_ctrlPts, BSpline.hpp:229-242). */
int bspl_template_function_from_ctrl(const bspl_template* t, const void* ctrl, int64_t n_fields,
                                     int on_device, void* stream, bspl_function** out);

/* BSpline(periodicity, ctrl_pts, knot_iter_pairs...) (BSpline.hpp:188-210): a
 * spline straight from knot vectors and control points (host pointers).
 * n_knots[d] - n_ctrl[d] must be order+1, or 2*order+1 on periodic axes. */
int bspl_function_from_control_points(bspl_dtype dtype, int dim, int order,
                                      const int64_t* n_ctrl, const int* periodic,
                                      const double* const* knots, const int64_t* n_knots,
                                      const void* ctrl, int64_t n_fields, int device,
                                      bspl_function** out);
/* copy constructor of InterpolationFunction (value semantics) */
int bspl_function_clone(const bspl_function* fn, bspl_function** out);
void bspl_function_destroy(bspl_function* fn);

/* periodicity(d), uniform(d), range(d), knots_num(d), get_order()
 * (Interpolation.hpp:248-267, BSpline.hpp:583-608).  Any out pointer may be
 * NULL.  n/periodic/uniform/n_knots/range_lo/range_hi have dim entries. */
int bspl_function_info(const bspl_function* fn, int* dtype, int* dim, int* order,
                       int64_t* n_fields, int64_t* n, int* periodic, int* uniform,
                       int64_t* n_knots, double* range_lo, double* range_hi);
/* CUDA ordinal of the device that holds the function's control points (no reference
 * counterpart: the reference's splines live in host memory). */
int bspl_function_device(const bspl_function* fn, int* device);
/* knots_begin(d)..knots_end(d) (BSpline.hpp:560-571), as double. */
int bspl_function_knots(const bspl_function* fn, int axis, double* out, int64_t capacity);
/* spline().control_points() in the plain (non-cell) layout (BSpline.hpp:575-577,
 * :45-50): copies field `field` to host memory, row-major, dtype of the function. */
int bspl_function_control_points(const bspl_function* fn, int64_t field, void* host_out);

/* operator()(DimArray<coord>) (Interpolation.hpp:132-146) when deriv == NULL,
 * derivative(coord, derivatives) (:177-205) otherwise (deriv has dim entries;
 * any entry > order yields 0, BSpline.hpp:404-407).  No bounds check:
 * out-of-range points extrapolate (non-periodic) or wrap (periodic).
 * pts [q][dim] -> out [q]. */
int bspl_evaluate(const bspl_function* fn, int64_t field, const void* pts, int64_t q,
                  const int* deriv, void* out, int on_device, void* stream);
/* at(coord) / derivative_at(coord, derivatives) (Interpolation.hpp:153-169,
 * :213-244): as bspl_evaluate but returns BSPL_ERR_DOMAIN, with *first_bad =
 * index of the first offending query, if a non-periodic coordinate lies outside
 * range(d) (boundary_check_, :482-491).  Host pointers only. */
int bspl_evaluate_at(const bspl_function* fn, int64_t field, const void* pts, int64_t q,
                     const int* deriv, void* out, int64_t* first_bad);
/* Fused value + gradient: out [q][1+dim] = {f, df/dx0, ..., df/dx(dim-1)}; what
 * the reference obtains from one operator() and dim derivative() calls.  A device
 * `out` for a 3-D function should be aligned to 4 elements (32 bytes for fp64):
 * the tiled kernel writes each result as one vector store; other alignments are
 * served by the (slower for large batches) direct kernel. */
int bspl_evaluate_value_grad(const bspl_function* fn, int64_t field, const void* pts, int64_t q,
                             void* out, int on_device, void* stream);
/* One query set applied to every field (the device analogue of eval_proxy,
 * InterpolationTemplate.hpp:145-176 / BSpline.hpp:244-297, following
 * operator() where the two disagree): out [n_fields][q]. */
int bspl_evaluate_fields(const bspl_function* fn, const void* pts, int64_t q, void* out,
                         int on_device, void* stream);
/* The same with query-major results, out [q][n_fields]: every query's values of all fields side by
 * side -- what the reference returns for a vector-valued T (one T{...} per query,
 * interpolation-test.cpp:674-703).  deriv as in bspl_evaluate (NULL: values).  This is the natural
 * layout of the kernel that serves many fields: queries sorted by cell, one small dense product
 * (queries of the cell) x (O+1)^D x (fields) per cell, control points read field-minor, whole lines
 * written per query.  A device `out` should be aligned to 4 elements and n_fields a multiple of 4;
 * other shapes are served by the field-major kernels plus a transpose. */
int bspl_evaluate_fields_query_major(const bspl_function* fn, const void* pts, int64_t q,
                                     const int* deriv, void* out, int on_device, void* stream);
/* eval_proxy (InterpolationTemplate.hpp:145-176, Interpolation.hpp:493-506): the work that
 * depends only on the query points -- locating them and, on the binned path, sorting them by
 * coefficient tile -- done once and reused for any number of evaluations (any field, any
 * derivative, any function built from the same template).  The plan keeps its own device copy
 * of what it needs; `pts` may be released after the call returns (host) / after the stream
 * reaches this point (device). */
typedef struct bspl_query_plan bspl_query_plan;
int bspl_query_plan_create(const bspl_function* fn, const void* pts, int64_t q, int on_device,
                           void* stream, bspl_query_plan** out);
/* InterpolationFunctionTemplate::eval_proxy (InterpolationTemplate.hpp:145-165): the same plan made
 * from the template alone, before any field has been interpolated -- a plan depends only on the
 * knots.  It serves every function this template produces. */
int bspl_template_query_plan_create(const bspl_template* t, const void* pts, int64_t q, int on_device,
                                    void* stream, bspl_query_plan** out);
/* value_grad == 0: out[q] (deriv == NULL: values); != 0: out[q][1+dim].  Results are in the
 * order of the original points.  A device `out` of a value+gradient evaluation through a tiled
 * (3-D, large-batch) plan must be aligned to 4 elements (32 bytes for fp64): the kernel stores
 * each result as one vector; any other alignment returns BSPL_ERR_INVALID. */
int bspl_query_plan_evaluate(const bspl_query_plan* plan, const bspl_function* fn, int64_t field,
                             const int* deriv, int value_grad, void* out, int on_device,
                             void* stream);
void bspl_query_plan_destroy(bspl_query_plan* plan);

/* get_knot_iter (BSpline.hpp:125-157): cell[q][dim] = span - order per axis,
 * exactly the first control-point index the reference selects. */
int bspl_locate(const bspl_function* fn, const void* pts, int64_t q, int32_t* cell,
                int on_device, void* stream);

/* Raw banded solver, BandLU<BandMatrix>/<ExtendedBandMatrix> (BandLU.hpp): factor
 * the dense row-major n x n matrix `a` (entries outside the band and the cyclic
 * corners are ignored) and solve n_rhs right-hand sides [n_rhs][n] in place on
 * the device.  Host pointers. */
int bspl_band_solve(int64_t n, int64_t p, int64_t q, int cyclic, const double* a, double* x,
                    int64_t n_rhs, int device);

/* The same solver fed with the band itself: rows[n][p+q+1], rows[i][k] = A(i, i + k - p); on a
 * cyclic matrix the column index wraps modulo n (the corner blocks of ExtendedBandMatrix,
 * BandMatrix.hpp:99-181), otherwise entries outside the matrix are ignored.  O(n (p+q)) memory. */
int bspl_band_solve_rows(int64_t n, int64_t p, int64_t q, int cyclic, const double* rows, double* x,
                         int64_t n_rhs, int device);

/* Host-only introspection of template construction (no device needed): the knot
 * vector of one axis exactly as create_knot_vector_ builds it, and the factored
 * collocation matrix in the row form the kernels consume.  factor arrays:
 * L[n][P], U[n][P], diag[n] with P = *band (= max(p, q)); for periodic axes the
 * corner strips bottom[n][P], right[n][P] (see bspl_kernels.h: AxisLU).  Pass NULL
 * to query sizes only.  coords as in bspl_template_create (NULL = uniform). */
int bspl_host_axis_knots(bspl_dtype dtype, int order, int periodic, int64_t n, double lo, double hi,
                         const double* coords, double* knots_out, int64_t capacity,
                         int64_t* n_knots, double* range_lo_hi);
int bspl_host_axis_factor(bspl_dtype dtype, int order, int periodic, int64_t n, double lo, double hi,
                          const double* coords, int* band, double* L, double* U, double* diag,
                          double* bottom, double* right);

/* Execution knobs.  path: 0 = auto, 1 = direct gather, 2 = cell-binned tiles. */
int bspl_set_eval_path(int path);
/* Many-field evaluation.  path: 0 = auto, 1 = per-query gather out of shared memory (field-major
 * kernel), 2 = cell-sorted contraction (query-major kernel; field-major results are transposed). */
int bspl_set_fields_path(int path);
/* Control-point solve.  path: 0 = auto, 1 = thread-per-line sweeps only (every line of the mesh in
 * flight), 2 = the tiled sweep that keeps the lines in flight inside the L2 (one read + one write of
 * the mesh per axis) wherever TMA can address the lines; auto takes it for meshes larger than the L2. */
int bspl_set_sweep_path(int path);
/* Number of kernels this library launched since the last reset (all threads). */
int64_t bspl_launch_count(void);
void bspl_reset_launch_count(void);
/* Device time (ms) of the most recent host-pointer call's kernels only (CUDA
 * events around the launches, excluding the copies); < 0 if none. */
double bspl_last_kernel_ms(void);

const char* bspl_last_error(void);
const char* bspl_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BSPLINE_B200_H */
