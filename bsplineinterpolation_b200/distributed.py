"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed for the
plumbing).  Only the single large 3-D solve has a real exchange step; everything else
shards into independent units:

  * evaluation      -- queries split evenly over ranks, coefficients replicated, no collective
                       (bench.py);
  * many fields     -- fields split over ranks, every rank holds the (tiny) LU factors (row 2);
  * one 3-D solve   -- slab-sharded along axis 0: sweep axes 2 and 1 locally, one all-to-all
                       re-shards to slabs along axis 1, sweep axis 0 (row 3).  The sweeps are
                       bspl_template_sweep_axis launches; the exchange is NCCL over NVLink.

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
    """A raw device allocation seen as a float64 CUDA array (plumbing: lets torch view memory that
    the library allocated for IPC)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": "<f8",
                                         "data": (int(ptr), False), "version": 3, "strides": None}

    def tensor(self, device):
        import torch
        return torch.as_tensor(self, device=torch.device("cuda", device))


class ShardedSolve3D:
    """Slab-sharded control-point solve of ONE 3-D field over the ranks of `group`."""

    def __init__(self, order, shape, ranges, periodicity=None, device=0, group=None):
        from .interpolation import InterpolationFunctionTemplate
        assert len(shape) == 3
        self.order, self.shape = int(order), tuple(int(s) for s in shape)
        self.periodicity = [bool(p) for p in (periodicity or [False] * 3)]
        self.group = group
        # every rank factors the three small collocation matrices itself
        self.template = InterpolationFunctionTemplate(order, shape, ranges, self.periodicity, device=device)

    def _shift(self, x, axis):
        import torch
        # periodic right-hand sides are rotated by O/2 (InterpolationTemplate.hpp:455-459)
        if self.periodicity[axis] and self.order // 2:
            return torch.roll(x, shifts=self.order // 2, dims=axis)
        return x

    def _local_sweeps(self, f_slab):
        """Working copy of the slab (periodic axes 1, 2 rotated) with axis 2 solved: lines along the
        contiguous axis go through the TMA-tiled sweep (bspl_solve.cu) in place."""
        n0, n1, n2 = self.shape
        w = self._shift(self._shift(f_slab, 2), 1)
        w = w.clone() if w.data_ptr() == f_slab.data_ptr() else w.contiguous()
        self.template.sweep_axis(2, w, (1, f_slab.shape[0], n1), (0, n1 * n2, n2), 1)
        return w

    def solve(self, f_slab, back_to_axis0=False):
        """f_slab: CUDA tensor [n0_loc, n1, n2], this rank's axis-0 slab of the mesh.
        Returns the control points as a slab of axis 1, [n0, n1_loc, n2] (or of axis 0)."""
        n0, n1, n2 = self.shape
        n0_loc = f_slab.shape[0]
        t = self.template
        w = self._local_sweeps(f_slab)
        t.sweep_axis(1, w, (1, n0_loc, n2), (0, n1 * n2, 1), n2)
        y = reshard_axis0_to_axis1(w, n0, self.group)
        y = self._shift(y, 0).contiguous()
        n1_loc = y.shape[1]
        t.sweep_axis(0, y, (1, 1, n1_loc * n2), (0, 0, 1), n1_loc * n2)
        return reshard_axis1_to_axis0(y, n1, self.group) if back_to_axis0 else y

    # ---- fused exchange: the last local sweep stores straight into the owners' buffers ----
    def enable_fused_exchange(self):
        """Allocate this rank's receive buffer [n0][n1_loc][n2] (bspl_ipc_alloc) and map every
        peer's buffer for access from this rank's GPU (bspl_ipc_open).  Collective."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        from ._capi import check, lib
        rank, world = _rank_world(self.group)
        n0, n1, n2 = self.shape
        s1 = shard_sizes(n1, world)
        dev = torch.cuda.current_device()
        self._dev = dev
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        check(lib().bspl_ipc_alloc(dev, n0 * s1[rank] * n2 * 8, C.byref(ptr), handle))
        self._own_ptr = ptr.value
        self._recv = _DevicePtr(ptr.value, (n0, s1[rank], n2)).tensor(dev)
        self._opened = []
        self._flag = torch.zeros(1, dtype=torch.float32, device="cuda")
        if world == 1:
            self._peers = [self._recv]
            return
        everyone = [None] * world
        dist.all_gather_object(everyone, bytes(handle), group=self.group)
        self._peers = []
        for r in range(world):
            if r == rank:
                self._peers.append(self._recv)
                continue
            h = (C.c_ubyte * 64).from_buffer_copy(everyone[r])
            p = C.c_void_p()
            check(lib().bspl_ipc_open(dev, h, C.byref(p)))
            self._opened.append(p.value)
            self._peers.append(_DevicePtr(p.value, (n0, s1[r], n2)).tensor(dev))

    def close_fused_exchange(self):
        """Unmap the peers' buffers, then (after a barrier) free this rank's.  Collective."""
        import torch
        import torch.distributed as dist
        from ._capi import check, lib
        torch.cuda.synchronize()
        self._peers = []
        for p in getattr(self, "_opened", []):
            check(lib().bspl_ipc_close(self._dev, p))
        self._opened = []
        if _rank_world(self.group)[1] > 1:
            dist.barrier(group=self.group)
        if getattr(self, "_own_ptr", None):
            self._recv = None
            check(lib().bspl_ipc_free(self._dev, self._own_ptr))
            self._own_ptr = None

    def solve_fused(self, f_slab):
        """As solve(), but the axis-1 sweep writes its solved rows directly into the receive buffers
        of their owners over NVLink (bspl_template_sweep_axis_exchange): no pack, no NCCL
        all-to-all, no unpack.  Returns this rank's slab of axis 1, [n0, n1_loc, n2]; the buffer
        is reused by the next call."""
        import torch
        import torch.distributed as dist
        rank, world = _rank_world(self.group)
        n0, n1, n2 = self.shape
        n0_loc = f_slab.shape[0]
        x0 = shard_range(n0, rank, world)[0]
        s1 = shard_sizes(n1, world)
        t = self.template
        w = self._local_sweeps(f_slab)
        if world > 1:
            # stream-ordered barrier (a one-element NCCL all-reduce, no host synchronisation): every
            # rank has consumed its previous result before anyone overwrites the buffers
            dist.all_reduce(self._flag, group=self.group)
        split = np.concatenate([[0], np.cumsum(s1)])
        t.sweep_axis_exchange(1, w, (1, n0_loc, n2), (0, n1 * n2, 1), n2, split,
                              [self._peers[r][x0:] for r in range(world)], [-1] * world,
                              [(0, s1[r] * n2, 1) for r in range(world)], [n2] * world)
        if world > 1:
            # every rank's exchange kernel has completed (its peer stores are visible at kernel end)
            dist.all_reduce(self._flag, group=self.group)
        y = self._recv
        if self.periodicity[0] and self.order // 2:
            y = self._shift(y, 0).contiguous()
        n1_loc = y.shape[1]
        t.sweep_axis(0, y, (1, 1, n1_loc * n2), (0, 0, 1), n1_loc * n2)
        return y

    def gather_function(self, ctrl_axis1_slab):
        """All-gather the solved slabs and build a replicated InterpolationFunction."""
        import torch
        import torch.distributed as dist
        world = _rank_world(self.group)[1]
        if world == 1:
            return self.template.function_from_control_points(ctrl_axis1_slab.contiguous())
        n0, n1, n2 = self.shape
        s1 = shard_sizes(n1, world)
        parts = [torch.empty((n0, s1[r], n2), dtype=ctrl_axis1_slab.dtype, device=ctrl_axis1_slab.device)
                 for r in range(world)]
        dist.all_gather(parts, ctrl_axis1_slab.contiguous(), group=self.group)
        full = torch.cat(parts, dim=1).contiguous()
        return self.template.function_from_control_points(full)
