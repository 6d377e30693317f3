"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the
plumbing).  Only the single large 3-D solve has a real exchange step; everything else
shards into independent units:

  * evaluation      -- queries split evenly over ranks, coefficients replicated, no collective
                       (bench.py);
  * many fields     -- fields split over ranks, every rank holds the (tiny) LU factors (row 2);
  * one 3-D solve   -- slab-sharded along axis 0: sweep axes 2 and 1 locally, one exchange
                       re-shards to slabs along axis 1, sweep axis 0 (row 3).  The whole sequence is
                       the library's bspl_sharded_solve_* plan: either one kernel that sweeps axis 1
                       and stores the solved rows into the peers' buffers over NVLink, or the same
                       kernel packing the blocks of an NCCL all-to-all.

The reference has no counterpart (single process, DedicatedThreadPool.hpp); the arithmetic
per line is that of solve_for_control_points_ (InterpolationTemplate.hpp:448-580), whose
per-axis solves commute across axes.
"""
import numpy as np


def shard_range(total, rank, world):
    """[begin, end) of `rank`'s share of `total` independent units (remainder to low ranks)."""
    base, rem = divmod(int(total), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_sizes(total, world):
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def _rank_world(group=None):
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _all_to_all(send, send_counts, recv_counts, group=None):
    import torch
    import torch.distributed as dist
    if _rank_world(group)[1] == 1:
        return send.clone()
    recv = torch.empty(int(sum(recv_counts)), dtype=send.dtype, device=send.device)
    try:
        dist.all_to_all_single(recv, send, output_split_sizes=[int(c) for c in recv_counts],
                               input_split_sizes=[int(c) for c in send_counts], group=group)
    except (RuntimeError, NotImplementedError):
        # backends without all_to_all (some gloo builds): pairwise exchange
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        so = np.concatenate([[0], np.cumsum(send_counts)]).astype(int)
        ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(int)
        reqs = []
        for r in range(world):
            if r == rank:
                recv[ro[r]:ro[r + 1]] = send[so[r]:so[r + 1]]
                continue
            reqs.append(dist.isend(send[so[r]:so[r + 1]].contiguous(), dst=r, group=group))
            reqs.append(dist.irecv(recv[ro[r]:ro[r + 1]], src=r, group=group))
        for q in reqs:
            q.wait()
    return recv


def reshard_axis0_to_axis1(local, n0, group=None):
    """local [n0_loc, n1, n2] (slab of axis 0) -> [n0, n1_loc, n2] (slab of axis 1)."""
    import torch
    rank, world = _rank_world(group)
    n0_loc, n1, n2 = local.shape
    s0, s1 = shard_sizes(n0, world), shard_sizes(n1, world)
    assert n0_loc == s0[rank]
    pieces, off = [], 0
    for r in range(world):
        pieces.append(local[:, off:off + s1[r], :].reshape(-1))
        off += s1[r]
    send = torch.cat(pieces)
    send_counts = [n0_loc * s1[r] * n2 for r in range(world)]
    recv_counts = [s0[r] * s1[rank] * n2 for r in range(world)]
    recv = _all_to_all(send, send_counts, recv_counts, group)
    # the block from rank r is [s0[r], n1_loc, n2] row-major: concatenation along axis 0 is free
    return recv.view(n0, s1[rank], n2)


def reshard_axis1_to_axis0(local, n1, group=None):
    """Inverse of reshard_axis0_to_axis1: [n0, n1_loc, n2] -> [n0_loc, n1, n2]."""
    import torch
    rank, world = _rank_world(group)
    n0, n1_loc, n2 = local.shape
    s0, s1 = shard_sizes(n0, world), shard_sizes(n1, world)
    assert n1_loc == s1[rank]
    send_counts = [s0[r] * n1_loc * n2 for r in range(world)]
    recv_counts = [s0[rank] * s1[r] * n2 for r in range(world)]
    recv = _all_to_all(local.reshape(-1), send_counts, recv_counts, group)
    out = torch.empty((s0[rank], n1, n2), dtype=local.dtype, device=local.device)
    off_e, off_y = 0, 0
    for r in range(world):
        cnt = recv_counts[r]
        out[:, off_y:off_y + s1[r], :] = recv[off_e:off_e + cnt].view(s0[rank], s1[r], n2)
        off_e += cnt
        off_y += s1[r]
    return out


class _DevicePtr:
    """A raw device allocation seen as a CUDA array (plumbing: lets torch view memory that the
    library owns)."""

    def __init__(self, ptr, shape, dtype=np.float64):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape),
                                         "typestr": "<f8" if np.dtype(dtype) == np.float64 else "<f4",
                                         "data": (int(ptr), False), "version": 3, "strides": None}

    def tensor(self, device):
        import torch
        return torch.as_tensor(self, device=torch.device("cuda", device))


class ShardedSolve3D:
    """Slab-sharded control-point solve of ONE 3-D field over the ranks of `group`: a thin holder of
    the library's plan (bspl_sharded_solve_*, include/bspline_b200.h).  The sweeps, the exchange and
    the rank barriers are launches of the C library on the current CUDA stream; torch.distributed
    only carries the 64-byte IPC handles at set-up and, for solve(), the NCCL all-to-all."""

    def __init__(self, order, shape, ranges, periodicity=None, device=0, group=None, dtype=np.float64):
        import ctypes as C
        import torch.distributed as dist
        from ._capi import check, lib
        from .interpolation import InterpolationFunctionTemplate
        assert len(shape) == 3
        self.order, self.shape = int(order), tuple(int(s) for s in shape)
        self.periodicity = [bool(p) for p in (periodicity or [False] * 3)]
        self.group = group
        self.device = int(device)
        self.dtype = np.dtype(dtype)
        self.rank, self.world = _rank_world(group)
        # every rank factors the three small collocation matrices itself
        self.template = InterpolationFunctionTemplate(order, shape, ranges, self.periodicity, dtype=dtype, device=device)
        h = C.c_void_p()
        check(lib().bspl_sharded_solve_create(self.template._h, self.rank, self.world, C.byref(h)))
        self._h = h
        b0 = (C.c_int64 * (self.world + 1))()
        b1 = (C.c_int64 * (self.world + 1))()
        check(lib().bspl_sharded_solve_layout(self._h, b0, b1))
        self.slab0, self.slab1 = list(b0), list(b1)
        # publish the receive buffers (CUDA IPC) and map the peers'
        handle = (C.c_ubyte * 64)()
        check(lib().bspl_sharded_solve_handle(self._h, handle))
        if self.world > 1:
            everyone = [None] * self.world
            dist.all_gather_object(everyone, bytes(handle), group=group)
            blob = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(everyone))
            check(lib().bspl_sharded_solve_connect(self._h, blob))
            dist.barrier(group=group)
        else:
            check(lib().bspl_sharded_solve_connect(self._h, None))

    def close(self):
        """Unmap the peers' buffers and free this rank's (collective: nobody may still be writing)."""
        import torch
        import torch.distributed as dist
        from ._capi import lib
        if getattr(self, "_h", None):
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)
            lib().bspl_sharded_solve_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            from ._capi import lib
            if getattr(self, "_h", None):
                lib().bspl_sharded_solve_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # kept for callers of the round-1 interface: the buffers now belong to the plan
    def enable_fused_exchange(self):
        pass

    def close_fused_exchange(self):
        pass

    def _slab_ptr(self, f_slab):
        import ctypes as C
        import torch
        want = torch.float64 if self.dtype == np.float64 else torch.float32
        n0, n1, n2 = self.shape
        exp = (self.slab0[self.rank + 1] - self.slab0[self.rank], n1, n2)
        if not (f_slab.is_cuda and f_slab.dtype == want and f_slab.is_contiguous() and tuple(f_slab.shape) == exp):
            raise ValueError("f_slab must be a contiguous %s CUDA tensor of shape %s" % (want, exp))
        if f_slab.device.index != self.device:
            raise ValueError("f_slab on %s, the plan on cuda:%d" % (f_slab.device, self.device))
        return C.c_void_p(f_slab.data_ptr()), C.c_void_p(torch.cuda.current_stream(f_slab.device).cuda_stream)

    def _view(self, ptr, shape):
        return _DevicePtr(ptr, shape, self.dtype).tensor(self.device)

    def solve_fused(self, f_slab):
        """One call, no host synchronisation: the axis-1 sweep stores its solved rows straight into
        their owners' buffers over NVLink (bspl_sharded_solve_run).  Returns this rank's slab of
        axis 1, [n0, n1_loc, n2]: a view of the plan's buffer, valid until the next solve."""
        import ctypes as C
        from ._capi import check, lib
        fp, sp = self._slab_ptr(f_slab)
        out = C.c_void_p()
        check(lib().bspl_sharded_solve_run(self._h, fp, C.byref(out), sp))
        n0, n1, n2 = self.shape
        return self._view(out.value, (n0, self.slab1[self.rank + 1] - self.slab1[self.rank], n2))

    def solve(self, f_slab, back_to_axis0=False):
        """The same solve with the exchange as an NCCL all-to-all: bspl_sharded_solve_pack (sweeps +
        packed send blocks), all_to_all_single, bspl_sharded_solve_finish (axis-0 sweep)."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from ._capi import check, lib
        fp, sp = self._slab_ptr(f_slab)
        send, recv, out = C.c_void_p(), C.c_void_p(), C.c_void_p()
        sc = (C.c_int64 * self.world)()
        rc = (C.c_int64 * self.world)()
        check(lib().bspl_sharded_solve_pack(self._h, fp, C.byref(send), C.byref(recv), sc, rc, sp))
        n0, n1, n2 = self.shape
        n1_loc = self.slab1[self.rank + 1] - self.slab1[self.rank]
        if self.world > 1:
            tsend = self._view(send.value, (int(sum(sc)),))
            trecv = self._view(recv.value, (int(sum(rc)),))
            dist.all_to_all_single(trecv, tsend, output_split_sizes=[int(c) for c in rc],
                                   input_split_sizes=[int(c) for c in sc], group=self.group)
        else:  # one rank: the "exchange" is a copy on the same stream
            self._view(recv.value, (int(sum(rc)),)).copy_(self._view(send.value, (int(sum(sc)),)))
        check(lib().bspl_sharded_solve_finish(self._h, C.byref(out), sp))
        y = self._view(out.value, (n0, n1_loc, n2))
        return reshard_axis1_to_axis0(y, n1, self.group) if back_to_axis0 else y

    def timed_out(self):
        import ctypes as C
        from ._capi import check, lib
        flag = C.c_int(0)
        check(lib().bspl_sharded_solve_status(self._h, C.byref(flag)))
        return bool(flag.value)

    def gather_function(self, ctrl_axis1_slab):
        """All-gather the solved slabs and build a replicated InterpolationFunction."""
        return self.template.function_from_control_points(self.gather_control_points(ctrl_axis1_slab))

    def gather_control_points(self, ctrl_axis1_slab):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return ctrl_axis1_slab.contiguous()
        n0, n1, n2 = self.shape
        parts = [torch.empty((n0, self.slab1[r + 1] - self.slab1[r], n2), dtype=ctrl_axis1_slab.dtype,
                             device=ctrl_axis1_slab.device) for r in range(self.world)]
        dist.all_gather(parts, ctrl_axis1_slab.contiguous(), group=self.group)
        return torch.cat(parts, dim=1).contiguous()
