#!/usr/bin/env python3
"""Headline benchmark: 3-D cubic fp64 value+gradient evaluation on a 256^3 mesh
(BASELINE.json configs[2]), one process per GPU, queries sharded across ranks with
replicated coefficients and no collective on the data path (weak scaling: every
rank evaluates Q_PER_GPU queries per step).  Also reports the second half of the
metric, the 512^3 control-point solve time, under "solve".

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One JSON line on stdout (rank 0).  `--impl reference` times the reference's own
CPU implementation (oracle/_ref, the unmodified headers compiled by
oracle/Makefile; falls back to the C port when that build is absent) on the
host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MESH = (256, 256, 256)
ORDER = 3
Q_PER_GPU = 1 << 28          # queries per rank per step (BASELINE configs[2]: 256M)
CPU_SAMPLE = 1 << 21         # queries timed on the CPU (value + 3 derivative calls each)
B_QUERY = 8 * 3 + 8 * 4 + 8 * 64   # algorithmic bytes per query, SURVEY 8(d): coords + out + stencil
B_STREAM = 8 * 3 + 8 * 4           # streamed bytes only
SOLVE_MESH = (512, 512, 512)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def smooth_field_np(shape):
    """prod_d cos(2 pi i_d / N_d - pi)  (interpolation-speed-test.cpp:84-89)"""
    f = np.ones(shape)
    for d, n in enumerate(shape):
        ax = np.cos(2 * np.pi * np.arange(n) / n - np.pi)
        f = f * ax.reshape([-1 if e == d else 1 for e in range(len(shape))])
    return f


# ----------------------------------------------------------------------------- reference arm
class CpuReference:
    """value + gradient through the reference: one operator() and three derivative() calls
    per point (the reference has no fused gradient), split over all host threads."""

    def __init__(self, threads):
        from oracle import pyoracle
        pyoracle.build()
        self.threads = threads
        f = smooth_field_np(MESH)
        t0 = time.perf_counter()
        if pyoracle.ref_available():
            self.sp = pyoracle.RefSpline(ORDER, f, [0, 0, 0], lo=[0, 0, 0], hi=[1, 1, 1], kind="cell")
            self.kind = "reference"
        else:
            self.sp = pyoracle.OracleSpline(ORDER, MESH, [0, 0, 0], lo=[0, 0, 0], hi=[1, 1, 1], f=f,
                                            nthreads=threads)
            self.kind = "port"
        self.construct_s = time.perf_counter() - t0
        self.pts = np.random.default_rng(12345).uniform(0.0, 1.0, (CPU_SAMPLE, 3))

    def step(self, n=None):
        pts = self.pts if n is None else self.pts[:n]
        t0 = time.perf_counter()
        self.sp.eval(pts, self.threads)
        for dv in ([1, 0, 0], [0, 1, 0], [0, 0, 1]):
            self.sp.deriv(pts, dv, self.threads)
        return time.perf_counter() - t0

    def describe(self, dt):
        return ("%d uniform-random queries per step, operator() + 3 derivative() calls each, std::thread "
                "split over %d threads (%.2f s per step; construct %.1f s not timed)"
                % (CPU_SAMPLE, self.threads, dt, self.construct_s))


def cpu_reference_solve():
    """The reference's own InterpolationFunctionTemplate::interpolate on the 512^3 mesh, with its
    INTP_MULTITHREAD pool (hard-wired to 8 workers + the caller, InterpolationTemplate.hpp:554) and
    without the cell-layout fill; template construction is not timed."""
    from oracle import pyoracle
    pyoracle.build()
    lib = os.path.join(pyoracle.OUT, "libintp_ref_plain_mt.so")
    if not os.path.exists(lib):
        return None
    f = smooth_field_np(SOLVE_MESH)
    sp = pyoracle.RefSpline(ORDER, f, [0, 0, 0], lo=[0, 0, 0], hi=[1, 1, 1], kind="plain_mt")
    return {"ms": sp.interpolate_ms, "kind": "reference", "threads": "reference pool: 8 workers + caller",
            "host_cores": os.cpu_count(), "config": "INTP_MULTITHREAD, plain layout (no INTP_CELL_LAYOUT)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ref = CpuReference(threads)
    for _ in range(min(args.warmup, 2)):
        ref.step(CPU_SAMPLE // 8)
    total = sum(ref.step() for _ in range(args.steps))
    value = CPU_SAMPLE * args.steps / total / 1e6
    line = {
        "impl": "reference", "metric": "3D cubic fp64 value+gradient eval", "value": value, "unit": "Mpts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg3: 3D cubic 256^3 mesh, value+gradient; CPU step = %d-query sample" % CPU_SAMPLE,
                   "mesh": list(MESH), "order": ORDER},
        "cpu_baseline": {"value": value, "unit": "Mpts/s", "cores": threads, "kind": ref.kind,
                         "sample": ref.describe(total / args.steps)},
        "e2e": {"value": value, "unit": "Mpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin=None, t_end=None):
        """Summarise the samples taken between the two wall-clock instants (all if None)."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.strip().split(", ") for r in open(self.tmp.name) if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t_begin is not None and not (t_begin - 0.05 <= ts <= t_end + 0.05):
                    continue
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def numa_nodes_visible():
    try:
        return len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
    except Exception:
        return None


def bind_to_gpu_numa_node(local):
    """Plumbing for the end-to-end leg: run this rank (and first-touch its pinned buffers) on the
    NUMA node its GPU hangs off.  Returns the node, or None when the platform does not say."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None



def smooth_field_slab(shape, b, e):
    """planes [b, e) of axis 0 of smooth_field_np(shape), without building the whole mesh"""
    axes = [np.cos(2 * np.pi * np.arange(n) / n - np.pi) for n in shape]
    return (axes[0][b:e, None, None] * axes[1][None, :, None]) * axes[2][None, None, :]


def timed_collective(torch, dist, world, fn, reps=5, warm=3):
    """median over `reps` of the max-over-ranks device time of fn() (CUDA events on the current
    stream, ranks aligned by a barrier before every repetition)"""
    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        sync()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(float(t.item()))
    return float(np.median(ts)), ts


def sharded_solve_leg(B, torch, dist, local, rank, world, single_ctrl_host):
    """BASELINE configs[3] as it is stated: the 512^3 solve slab-sharded over the ranks, both
    exchanges (NCCL all-to-all; sweep fused with NVLink peer stores), and the gathered control
    points compared bit for bit with the single-GPU solve of rank 0."""
    from bsplineinterpolation_b200.distributed import ShardedSolve3D
    n0 = SOLVE_MESH[0]
    sh = ShardedSolve3D(ORDER, SOLVE_MESH, [(0.0, 1.0)] * 3, device=local)
    b, e = sh.slab0[rank], sh.slab0[rank + 1]
    f = torch.from_numpy(smooth_field_slab(SOLVE_MESH, b, e)).to(torch.device("cuda", local))
    res = {"mesh": list(SOLVE_MESH), "n_gpus": world, "planes_per_gpu": e - b,
           "exchange_bytes_per_gpu": int((e - b) * SOLVE_MESH[1] * SOLVE_MESH[2] * 8 * (world - 1) / world)}
    parity = {}
    for name, run in (("nccl", lambda: sh.solve(f)), ("fused", lambda: sh.solve_fused(f))):
        ms, all_ms = timed_collective(torch, dist, world, run)
        res[name + "_ms"] = ms
        res[name + "_ms_all"] = all_ms
        full = sh.gather_control_points(run())
        torch.cuda.synchronize()
        if rank == 0 and single_ctrl_host is not None:
            got = full.cpu().numpy()
            parity[name] = bool(np.array_equal(got, single_ctrl_host))
            if not parity[name]:   # where: a diagnosis travels with the line
                bad = np.argwhere(got != single_ctrl_host)
                res[name + "_mismatch"] = {"count": int(len(bad)), "first": bad[0].tolist(), "min": bad.min(0).tolist(),
                                           "max": bad.max(0).tolist(),
                                           "max_abs": float(np.abs(got - single_ctrl_host).max()),
                                           "planes0": np.unique(bad[:, 0])[:16].tolist(),
                                           "rows1": np.unique(bad[:, 1])[:16].tolist()}
            del got
        del full
        torch.cuda.empty_cache()
    if sh.timed_out():
        res["error"] = "a rank barrier timed out"
    res["parity"] = (all(parity.values()) if parity else None)
    res["parity_detail"] = parity
    res["parity_against"] = "control points of the single-GPU solve on rank 0, numpy.array_equal (bit-identical)"
    sbytes = 2 * 8 * 3 * float(np.prod(SOLVE_MESH))
    res["best_ms"] = min(res["nccl_ms"], res["fused_ms"])
    res["algorithmic_gbs_aggregate"] = sbytes / res["best_ms"] / 1e6
    sh.close()
    del f
    torch.cuda.empty_cache()
    return res


def sharded_fields_leg(B, torch, dist, local, rank, world):
    """BASELINE configs[4] field-sharded: the 4 096 fields split over the ranks,
    every rank holds the template, solves its fields and evaluates them at the same 2^20 points.
    No collective on the data path."""
    from bsplineinterpolation_b200.distributed import shard_range
    dev = torch.device("cuda", local)
    F, shape, Q = 4096, (128, 128), 1 << 20
    fb, fe = shard_range(F, rank, world)
    t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 2, device=local)
    gen = torch.Generator(device=dev)
    gen.manual_seed(777)                      # every rank draws the whole batch, keeps its share
    f = torch.rand((F,) + shape, dtype=torch.float64, device=dev, generator=gen)[fb:fe].contiguous()
    pts = torch.rand((Q, 2), dtype=torch.float64, device=dev, generator=gen)
    fn = t.interpolate(f)
    solve_ms, _ = timed_collective(torch, dist, world, lambda: t.interpolate(f, into=fn), reps=5, warm=1)
    out = torch.empty((fe - fb, Q), dtype=torch.float64, device=dev)
    eval_ms, _ = timed_collective(torch, dist, world, lambda: fn.evaluate_fields(pts, out=out), reps=3, warm=1)
    res = {"fields_total": F, "fields_per_gpu": fe - fb, "queries": Q, "solve_ms": solve_ms,
           "evaluate_fields_ms": eval_ms, "G_evaluations_per_s": F * Q / eval_ms / 1e6}
    try:
        outT = torch.empty((Q, fe - fb), dtype=torch.float64, device=dev)
        ms2, _ = timed_collective(torch, dist, world, lambda: fn.evaluate_fields(pts, out=outT, layout="query_major"),
                                  reps=3, warm=1)
        res["evaluate_fields_query_major_ms"] = ms2
        res["G_evaluations_per_s_query_major"] = F * Q / ms2 / 1e6
        k = (fe - fb) // 2
        one = fn.evaluate(pts, field=k)
        scale = float(one.abs().max().item())
        res["max_rel_diff_vs_per_field"] = max(float((out[k] - one).abs().max().item()),
                                               float((outT[:, k] - one).abs().max().item())) / scale
        del outT
    except Exception as exc:
        res["query_major_error"] = str(exc)
    del f, pts, out, fn, t
    torch.cuda.empty_cache()
    return res


def copy_ceiling(torch, dist, world, dev, h_pts, h_out, steps):
    """The end-to-end leg's copies with no kernel at all: the same chunking (2^22 queries), three
    streams, H2D of the points and D2H of a result-sized device buffer -- what the host<->device
    path of THIS box allows at THIS number of ranks."""
    qe = h_pts.shape[0]
    chunk = min(qe, 1 << 22)
    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
    d_in = [torch.empty((chunk, 3), dtype=torch.float64, device=dev) for _ in range(3)]
    d_out = [torch.zeros((chunk, 4), dtype=torch.float64, device=dev) for _ in range(3)]

    def one_pass():
        for it, lo in enumerate(range(0, qe, chunk)):
            hi = min(qe, lo + chunk)
            sl = it % 3
            with torch.cuda.stream(streams[sl]):
                d_in[sl][: hi - lo].copy_(h_pts[lo:hi], non_blocking=True)
                h_out[lo:hi].copy_(d_out[sl][: hi - lo], non_blocking=True)
        torch.cuda.synchronize()

    one_pass()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_pass()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return world * qe * steps / float(t.item()) / 1e6

# ----------------------------------------------------------------------------- other configurations
def other_configs(B, torch, dev, hbm_peak):
    """BASELINE.json configs 0, 1 and 4 (cfg1, cfg2, cfg5 in SURVEY 8), device resident, CUDA events,
    median of 5 after 2 warm-up calls; algorithmic bytes per SURVEY 8(d)."""
    def timed(fn, reps=5, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    def device_field(shape):
        return torch.from_numpy(smooth_field_np(shape)).to(dev)

    res = {}
    # cfg1: 2-D cubic 1024^2, value+gradient (the mesh is L2 resident: direct gather kernel)
    try:
        shape, Q = (1024, 1024), 1 << 24
        fn = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 2, device=dev.index).interpolate(device_field(shape))
        pts = torch.rand((Q, 2), dtype=torch.float64, device=dev)
        out = torch.empty((Q, 3), dtype=torch.float64, device=dev)
        ms = timed(lambda: fn.value_grad(pts, out=out))
        bq = 8 * 2 + 8 * 3 + 8 * 16
        res["cfg1_2d_cubic_1024"] = {"queries": Q, "outputs": "value+gradient", "ms": ms, "Mpts_per_s": Q / ms / 1e3,
                                      "bytes_per_query": bq, "roofline_frac": Q * bq / ms / 1e6 / hbm_peak}
        del fn, pts, out
    except Exception as exc:
        res["cfg1_2d_cubic_1024"] = {"error": str(exc)}
    torch.cuda.empty_cache()
    # cfg2: 1-D order-5 periodic, 2^24 mesh points, value + first derivative; solve time beside it
    try:
        shape, Q = (1 << 24,), 1 << 26   # BASELINE configs[1]: 64M queries
        t0 = time.perf_counter()
        t = B.InterpolationFunctionTemplate(5, shape, [(0.0, 1.0)], [True], device=dev.index)
        template_ms = 1e3 * (time.perf_counter() - t0)
        f = device_field(shape)
        fn = t.interpolate(f)
        solve_ms = timed(lambda: t.interpolate(f, into=fn), warm=1)
        pts = torch.rand((Q, 1), dtype=torch.float64, device=dev)
        out = torch.empty((Q, 2), dtype=torch.float64, device=dev)
        ms = timed(lambda: fn.value_grad(pts, out=out))
        bq = 8 + 8 * 2 + 8 * 6
        res["cfg2_1d_quintic_periodic_2e24"] = {"queries": Q, "outputs": "value+d1", "ms": ms, "Mpts_per_s": Q / ms / 1e3,
                                                 "bytes_per_query": bq, "roofline_frac": Q * bq / ms / 1e6 / hbm_peak,
                                                 "solve_ms": solve_ms, "template_host_ms": template_ms}
        del fn, pts, out, f, t
    except Exception as exc:
        res["cfg2_1d_quintic_periodic_2e24"] = {"error": str(exc)}
    torch.cuda.empty_cache()
    # cfg5: 4 096 fields on one 128^2 cubic mesh: batched solve, then one query set on every field
    try:
        F, shape, Q = 4096, (128, 128), 1 << 20
        t = B.InterpolationFunctionTemplate(3, shape, [(0.0, 1.0)] * 2, device=dev.index)
        f = torch.rand((F,) + shape, dtype=torch.float64, device=dev)
        fn = t.interpolate(f)
        solve_ms = timed(lambda: t.interpolate(f, into=fn), warm=1)
        pts = torch.rand((Q, 2), dtype=torch.float64, device=dev)
        out = torch.empty((F, Q), dtype=torch.float64, device=dev)
        ms = timed(lambda: fn.evaluate_fields(pts, out=out), reps=3, warm=1)
        sbytes = 2 * 8 * 2 * float(F * shape[0] * shape[1])
        res["cfg5_4096_fields_128x128"] = {"fields": F, "queries": Q, "solve_ms": solve_ms,
                                            "solve_roofline_frac": sbytes / solve_ms / 1e6 / hbm_peak,
                                            "evaluate_fields_ms": ms, "G_evaluations_per_s": F * Q / ms / 1e6,
                                            "output_stream_frac": F * Q * 8 / ms / 1e6 / hbm_peak}
        del out
        torch.cuda.empty_cache()
        # the same evaluation with query-major results, out[q][field] -- the layout the cell-sorted
        # contraction produces natively (bspl_evaluate_fields_query_major)
        try:
            outT = torch.empty((Q, F), dtype=torch.float64, device=dev)
            ms2 = timed(lambda: fn.evaluate_fields(pts, out=outT, layout="query_major"), reps=3, warm=1)
            k = F // 2
            one = fn.evaluate(pts, field=k)
            c5 = res["cfg5_4096_fields_128x128"]
            c5["evaluate_fields_query_major_ms"] = ms2
            c5["G_evaluations_per_s_query_major"] = F * Q / ms2 / 1e6
            c5["output_stream_frac_query_major"] = F * Q * 8 / ms2 / 1e6 / hbm_peak
            c5["max_rel_diff_vs_per_field"] = float((outT[:, k] - one).abs().max().item()) / float(one.abs().max().item())
            del outT, one
        except Exception as exc:
            res["cfg5_4096_fields_128x128"]["query_major_error"] = str(exc)
        del fn, pts, f, t
    except Exception as exc:
        res["cfg5_4096_fields_128x128"] = {"error": str(exc)}
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    import bsplineinterpolation_b200 as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    hbm_peak, peak_kind = peaks()
    q = args.queries
    # replicated coefficients: every rank solves the same 256^3 field (outside the timed region)
    tmpl = B.InterpolationFunctionTemplate(ORDER, MESH, [(0.0, 1.0)] * 3, device=local)
    f = torch.from_numpy(smooth_field_np(MESH)).to(dev)
    fn = tmpl.interpolate(f)
    del f
    gen = torch.Generator(device=dev)
    gen.manual_seed(12345 + rank)
    pts = torch.rand((q, 3), dtype=torch.float64, device=dev, generator=gen)
    out = torch.empty((q, 4), dtype=torch.float64, device=dev)

    def step():
        fn.value_grad(pts, out=out)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    t_begin = time.time()
    B.reset_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        ev[k][0].record()
        step()
        ev[k][1].record()
    e1.record()
    barrier()
    t_end = time.time()
    launches = B.launch_count()
    clocks = sampler.stop(t_begin, t_end)
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    value = world * q * args.steps / (total_ms * 1e-3) / 1e6
    achieved = q * B_QUERY / (kern_ms * 1e-3) / 1e9
    checksum = float(out[:: max(1, q // 4096)].sum().item())

    # ---- strong scaling companion (BASELINE configs[2] as stated: 2^28 queries SPLIT over the GPUs)
    strong = None
    if world > 1:
        qs = max(1, args.queries // world)
        s_pts, s_out = pts[:qs], out[:qs]
        for _ in range(2):
            fn.value_grad(s_pts, out=s_out)
        barrier()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            fn.value_grad(s_pts, out=s_out)
        b.record()
        barrier()
        s_ms = max_over_ranks(a.elapsed_time(b)) / args.steps
        strong = {"queries_total": qs * world, "queries_per_gpu": qs, "ms_per_step": s_ms,
                  "value": qs * world / s_ms / 1e3, "unit": "Mpts/s", "scaling": "strong",
                  "roofline_frac_per_gpu": qs * B_QUERY / (s_ms * 1e-3) / 1e9 / hbm_peak}

    # ---- the dominant kernel alone (eval_binned_kernel: the tile evaluation): a query plan keeps the
    # sorted records, so evaluating through it launches that kernel and nothing else
    dominant = None
    try:
        plan = fn.eval_proxy(pts)
        plan(fn, value_grad=True, out=out)
        torch.cuda.synchronize()
        B.reset_launch_count()
        pe = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        for a, b in pe:
            a.record(); plan(fn, value_grad=True, out=out); b.record()
        torch.cuda.synchronize()
        n_launch = max(1, B.launch_count() // len(pe))
        dom_ms = float(np.mean([a.elapsed_time(b) for a, b in pe]))
        dominant = {"name": "eval_binned_kernel<double,3,grad>", "launches_per_step": int(n_launch),
                    "ms_per_launch": dom_ms / n_launch, "share_of_step": dom_ms / kern_ms,
                    "achieved": q * B_QUERY / (dom_ms * 1e-3) / 1e9, "unit": "GB/s",
                    "note": "algorithmic 568 B/query; the 512 stencil bytes are served by shared memory, "
                            "so this exceeds the HBM peak; its floor is one random 32-byte result write per query"}
        del plan
        torch.cuda.empty_cache()
    except Exception as exc:  # the plan needs 32 B of scratch per query
        dominant = {"error": str(exc)}

    # ---- end to end: host buffers through the C ABI, copies inside the timed region
    qe = q if args.e2e_queries <= 0 else min(q, args.e2e_queries)
    try:
        import psutil
        avail = psutil.virtual_memory().available
        while qe > (1 << 22) and qe * 56 * world * 2 > avail:  # keep pinned buffers well inside host RAM
            qe //= 2
    except Exception:
        pass
    h_pts = torch.empty((qe, 3), dtype=torch.float64, pin_memory=True)
    h_out = torch.empty((qe, 4), dtype=torch.float64, pin_memory=True)
    h_pts.copy_(pts[:qe])
    np_pts, np_out = h_pts.numpy(), h_out.numpy()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    fn.value_grad(np_pts, out=np_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fn.value_grad(np_pts, out=np_out)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * qe * e2e_steps / e2e_s / 1e6
    e2e_err = float(np.abs(np_out[:4096] - out[:4096].cpu().numpy()).max())
    try:
        ceiling = copy_ceiling(torch, dist, world, dev, h_pts, h_out, e2e_steps)
    except Exception as exc:
        ceiling = None
        print("copy ceiling failed: %s" % exc, file=sys.stderr)
    del h_pts, h_out, np_pts, np_out

    # ---- second half of the metric: 512^3 control-point solve (device resident), rank-local
    solve = None
    if not args.no_solve:
        del pts, out
        torch.cuda.empty_cache()
        st = B.InterpolationFunctionTemplate(ORDER, SOLVE_MESH, [(0.0, 1.0)] * 3, device=local)
        sf = torch.from_numpy(smooth_field_np(SOLVE_MESH)).to(dev)
        sfn = st.interpolate(sf)
        for _ in range(2):
            st.interpolate(sf, into=sfn)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); st.interpolate(sf, into=sfn); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        sbytes = 2 * 8 * 3 * float(np.prod(SOLVE_MESH))
        solve = {"mesh": list(SOLVE_MESH), "ms": ms, "target_ms": 50.0,
                 "algorithmic_gbs": sbytes / ms / 1e6, "roofline_frac": sbytes / ms / 1e6 / hbm_peak}
        # the same solve end to end: the mesh starts in (pinned) host memory, as with the header's
        # interpolate(Mesh); the control points stay on the device, where evaluation needs them
        try:
            h_mesh = torch.empty(SOLVE_MESH, dtype=torch.float64, pin_memory=True)
            h_mesh.copy_(sf)
            np_mesh = h_mesh.numpy()
            st.interpolate(np_mesh, into=sfn)
            te = []
            for _ in range(3):
                t0 = time.perf_counter()
                st.interpolate(np_mesh, into=sfn)   # returns when the control points are in place
                te.append(1e3 * (time.perf_counter() - t0))
            solve["e2e_ms"] = float(np.median(te))
            solve["e2e_h2d_bytes"] = int(np.prod(SOLVE_MESH)) * 8
            del h_mesh, np_mesh
        except Exception as exc:
            solve["e2e_ms"] = None
            solve["e2e_error"] = str(exc)
        single_ctrl = None
        if world > 1 and rank == 0:
            single_ctrl = sfn.control_points()     # host copy, for the sharded solve's parity check
        del sf, sfn, st
        torch.cuda.empty_cache()
        if world > 1:
            try:
                solve["sharded"] = sharded_solve_leg(B, torch, dist, local, rank, world, single_ctrl)
            except Exception as exc:
                solve["sharded"] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            del single_ctrl
        if rank == 0 and world == 1 and not args.no_cpu:
            solve["cpu_reference"] = cpu_reference_solve()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        ref = CpuReference(threads)
        ref.step(CPU_SAMPLE // 8)
        dt = ref.step()
        cpu = {"value": CPU_SAMPLE / dt / 1e6, "unit": "Mpts/s", "cores": threads, "kind": ref.kind,
               "sample": ref.describe(dt)}

    # ---- the other BASELINE configurations, device resident (one GPU only; each guarded: a failure
    # is recorded, it cannot take the headline numbers above with it)
    other = None
    if world == 1 and not args.no_other:
        other = other_configs(B, torch, dev, hbm_peak)
    fields_sharded = None
    if world > 1 and not args.no_other:
        try:
            fields_sharded = sharded_fields_leg(B, torch, dist, local, rank, world)
        except Exception as exc:
            fields_sharded = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if rank == 0:
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                traffic = tj.get("eval_bytes_per_query")
                traffic = traffic * q if traffic is not None else None
                traffic_src = ("ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of key_count + scatter + eval_binned, "
                               "%d queries, commit %s (profiles/traffic.json), scaled to this run's query count"
                               % (tj.get("queries_in_capture", 0), tj.get("commit", "?")))
            except Exception:
                traffic = None
        line = {
            "metric": "3D cubic fp64 value+gradient eval", "value": value, "unit": "Mpts/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg3: 3D cubic 256^3 mesh, %d uniform-random queries per GPU per step, "
                                   "value+gradient, coefficients replicated, queries sharded" % q,
                       "mesh": list(MESH), "order": ORDER, "queries_per_gpu": q,
                       "l2": "inputs+outputs (%.1f GB per step) exceed the 126 MB L2" % (q * B_STREAM / 1e9),
                       "strong": strong},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
                         "kernel_ms": kern_ms, "kernel": "whole evaluate call (key_count, scans, scatter, eval_binned)",
                         "dominant_kernel": dominant, "bytes_per_query": B_QUERY,
                         "stream_only_frac": q * B_STREAM / (kern_ms * 1e-3) / 1e9 / hbm_peak,
                         # measured DRAM bytes (ncu, profiles/traffic.json) over the live step time
                         "dram_frac": (traffic / (kern_ms * 1e-3) / 1e9 / hbm_peak) if traffic else None},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "Mpts/s", "h2d_bytes_per_step": qe * 24, "d2h_bytes_per_step": qe * 32,
                    "queries_per_gpu": qe, "steps": e2e_steps, "max_abs_diff_vs_device_path": e2e_err,
                    "copy_ceiling": ceiling,
                    "frac_of_ceiling": (e2e_value / ceiling) if ceiling else None,
                    "copy_ceiling_note": "the same host<->device copies (2^22-query chunks, 3 streams, both directions) "
                                         "with no kernel, all ranks at once, max over ranks",
                    "numa_node_rank0": numa_node, "numa_nodes_visible": numa_nodes_visible(),
                    "host_cores": os.cpu_count()},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "solve": solve,
            "other_configs": other,
            "cfg5_field_sharded": fields_sharded,
            "checksum": checksum,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--queries", type=int, default=Q_PER_GPU, help="queries per GPU per step")
    ap.add_argument("--e2e-queries", type=int, default=0, help="0 = same as --queries")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the cfg1 / cfg2 / cfg5 timings")
    args = ap.parse_args()
    # stdout carries the one JSON line and nothing else: whatever libraries print while the run is
    # in progress (NCCL's version banner, compiler chatter of the CPU checker) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


_REAL_STDOUT = None


def emit(line):
    """Print the result line on the process's original stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
