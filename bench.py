#!/usr/bin/env python
"""bench.py — LETKF analysed grid points / s on B200 (BASELINE.json metric), with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg1|small] [--impl reference]

One "step" is one full LETKF analysis of the workload: observation binning + fused analysis kernel (+ for
N > 1 the broadcast of the observation-space arrays from rank 0 and the all-gather of the analysis).
`value` is measured with every input already resident in HBM; `e2e` goes through the host-buffer entry point
(`b200da_letkf_host`: pinned host arrays in, host array out, copies inside the timed region).
`--impl reference` times the reference's CPU algorithm (the numpy/LAPACK oracle port, oracle/letkf_oracle.py —
the reference itself is pure Python and cannot be imported without xarray/dask) on all host cores on a bounded
sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))

WORKLOADS = {
    # name: (description, generator kwargs, metric ctor, GC radius, inflation)
    "cfg3": dict(desc="cfg3: synthetic global 1000x1000 lat-lon grid (1M points), k=50, 2.5M obs uniform on the sphere, "
                      "haversine GC c=1000 km, inf_factor 1.1, FP64",
                 kind="sphere", nlat=1000, nlon=1000, k=50, n_obs=2_500_000, radius=1000.0, rho=1.1),
    "cfg2": dict(desc="cfg2: Lorenz-96 ring N=100k, k=40, every 2nd variable observed, periodic GC c=20, inf_factor 1.1, FP64",
                 kind="ring", n_grid=100_000, k=40, stride=2, radius=20.0, rho=1.1),
    "cfg1": dict(desc="cfg1: Lorenz-96 ring N=40, k=50, all observed, periodic GC c=5, inf_factor 1.1, FP64",
                 kind="ring", n_grid=40, k=50, stride=1, radius=5.0, rho=1.1),
    "small": dict(desc="small: 100x200 lat-lon grid, k=50, 50k obs, haversine GC c=1000 km (smoke-sized cfg3)",
                  kind="sphere", nlat=100, nlon=200, k=50, n_obs=50_000, radius=1000.0, rho=1.1),
}


def make_workload(name):
    from pytassim_b200.testing import synthetic as syn
    w = WORKLOADS[name]
    if w["kind"] == "sphere":
        data = syn.sphere_latlon(w["nlat"], w["nlon"], w["k"], w["n_obs"], seed=42)
    else:
        data = syn.lorenz96_1d(w["n_grid"], w["k"], w["stride"], seed=42)
    return w, data


def make_metric(w, data):
    from pytassim_b200.localization import metrics
    if w["kind"] == "sphere":
        return metrics.HaversineDistance(6371.0)
    return metrics.PeriodicDistance1D(data["period"])


def oracle_dist(w, data):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import letkf_oracle as orc
    if w["kind"] == "sphere":
        return orc, orc.make_dist_haversine(6371.0)
    return orc, orc.make_dist_periodic1d(data["period"])


# ---- CPU baseline: the reference algorithm (oracle port) on the host cores -----------------------------------
_G = {}


def _cpu_init():
    try:
        from threadpoolctl import threadpool_limits
        _G["limit"] = threadpool_limits(limits=1)       # reference advice: OMP_NUM_THREADS=1 per worker
    except Exception:
        pass


def _cpu_chunk(sel):
    orc, dist, w, data = _G["orc"], _G["dist"], _G["w"], _G["data"]
    ana, _ = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                data["obs_rows"], dist, w["radius"], inf_factor=w["rho"], grid_subset=sel)
    return float(ana.sum())


def cpu_reference_rate(w, data, target_seconds=15.0, max_points=None):
    """Grid points / s of the reference algorithm on all host cores; bounded sample of evenly spaced grid points."""
    import multiprocessing as mp
    orc, dist = oracle_dist(w, data)
    _G.update(orc=orc, dist=dist, w=w, data=data)
    cores = len(os.sched_getaffinity(0))
    n_grid = data["state"].shape[-1]
    # calibrate on a few points in this process
    probe = np.linspace(0, n_grid - 1, 2, dtype=np.int64)
    t0 = time.perf_counter(); _cpu_chunk(probe[:1]); _cpu_chunk(probe[1:]); per_point = (time.perf_counter() - t0) / 2
    n_pts = int(min(n_grid, max(cores, target_seconds * cores / max(per_point, 1e-6))))
    if max_points:
        n_pts = min(n_pts, max_points)
    sel = np.unique(np.linspace(0, n_grid - 1, n_pts, dtype=np.int64))
    chunks = [c for c in np.array_split(sel, cores * 4) if len(c)]
    # a workload smaller than the sample budget (cfg1: 40 grid points) is analysed several times over, so that the timed
    # region is CPU work and not the hand-over of 40 tasks to the pool
    est = per_point * len(sel) / cores
    passes = int(min(1000, max(1, 0.25 * target_seconds / max(est, 1e-6)))) if len(sel) == n_grid else 1
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        pool.map(_cpu_chunk, [c[:1] for c in chunks[:cores]])          # warm the workers
        t0 = time.perf_counter()
        pool.map(_cpu_chunk, chunks * passes)
        dt = time.perf_counter() - t0
    return dict(value=len(sel) * passes / dt, unit="gridpoints/s", cores=cores, kind="port",
                sample="{0} evenly spaced grid points of {1}{4} on {2} processes x 1 thread, numpy+LAPACK port of the "
                       "reference loop (localize_obs over all obs -> sqrt(w) gather -> ETKF weights -> update), "
                       "{3:.1f} s wall".format(len(sel), n_grid, cores, dt, " x {0} passes".format(passes) if passes > 1 else ""))


# ---- clocks ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        smax = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm))


def measure_fp64_peak(torch, seconds=1.5):
    """cuBLAS DGEMM 8192^3: burst (best of 5) and sustained (back to back) — the FP64 roofline denominator
    (same method as MEASURED_PEAKS.json, which has no FP64 entry)."""
    n = 8192
    a = torch.randn(n, n, device="cuda", dtype=torch.float64)
    b = torch.randn(n, n, device="cuda", dtype=torch.float64)
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    reps = max(3, int(seconds * 1000.0 / best))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); e1.synchronize()
    flops = 2.0 * n ** 3
    del a, b, c
    torch.cuda.empty_cache()
    return flops / best * 1e-9, flops * reps / e0.elapsed_time(e1) * 1e-9


_JSON_FD = None


def _reserve_stdout():
    """Keep the real stdout for the ONE JSON line and send everything else that writes to fd 1 (the NCCL version banner,
    library chatter of child processes) to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    payload = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(payload.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, payload)


def main():
    _reserve_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="plan dtype (state / obs-space arrays in HBM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric_name, unit = "letkf_analysed_gridpoints_per_sec", "gridpoints/s"

    # ------------------------------------------------------------------------------------------------------------
    # reference arm: the reference's CPU algorithm on the host cores (rank 0 only)
    # ------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        w, data = make_workload(args.workload)
        vals, info = [], None
        for _ in range(max(1, args.warmup > 0)):
            cpu_reference_rate(w, data, target_seconds=3.0)
        t_all = time.perf_counter()
        for _ in range(args.steps):
            info = cpu_reference_rate(w, data, target_seconds=12.0)
            vals.append(info["value"])
        v = float(np.mean(vals))
        info["value"] = v
        line = {"metric": metric_name, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1000.0 * (time.perf_counter() - t_all) / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "impl": "reference", "config": {"workload": w["desc"], "inputs": "host resident"},
                "cpu_baseline": info,
                "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        _emit(line)
        return

    # ------------------------------------------------------------------------------------------------------------
    # B200 arm
    # ------------------------------------------------------------------------------------------------------------
    w, data = (None, None)
    if rank == 0:
        w, data = make_workload(args.workload)
    else:
        w = WORKLOADS[args.workload]
    cpu_info = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_info = cpu_reference_rate(w, data, target_seconds=12.0)      # before CUDA is initialised (fork safety)

    import torch
    import torch.distributed as dist
    from pytassim_b200.engine import LETKFEngine, launch_count
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL writes its version banner to STDOUT when NCCL_DEBUG is VERSION: keep stdout to the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    k = w["k"]
    # shapes travel from rank 0
    if world > 1:
        shp = torch.zeros(4, dtype=torch.int64, device=dev)
        if rank == 0:
            shp[:] = torch.tensor([data["state"].shape[-1], data["normed_obs"].shape[0], data["grid_rows"].shape[1] - 1, 0])
        dist.broadcast(shp, 0)
        n_grid, n_obs, n_coord = int(shp[0]), int(shp[1]), int(shp[2])
    else:
        n_grid, n_obs, n_coord = data["state"].shape[-1], data["normed_obs"].shape[0], data["grid_rows"].shape[1] - 1

    f64 = torch.float64
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    ndt = np.float64 if args.dtype == "f64" else np.float32
    esz = 8 if args.dtype == "f64" else 4
    if rank == 0:
        x_host = torch.from_numpy(np.ascontiguousarray(data["state"].reshape(1, k, n_grid), dtype=ndt)).pin_memory()
        y_host = torch.from_numpy(np.ascontiguousarray(data["normed_perts"], dtype=ndt)).pin_memory()
        d_host = torch.from_numpy(np.ascontiguousarray(data["normed_obs"], dtype=ndt)).pin_memory()
        oc_host = torch.from_numpy(np.ascontiguousarray(data["obs_rows"][:, 1:].T)).pin_memory()     # (n_coord, M)
        gc_dev = torch.from_numpy(np.ascontiguousarray(data["grid_rows"][:, 1:])).to(dev)
        x_dev, y_dev, d_dev = x_host.to(dev), y_host.to(dev), d_host.to(dev)
        oc_dev = oc_host.to(dev).t().contiguous()                                                     # (M, n_coord)
    else:
        gc_dev = torch.empty((n_grid, n_coord), dtype=f64, device=dev)
        x_dev = torch.empty((1, k, n_grid), dtype=tdt, device=dev)
        y_dev = torch.empty((k, n_obs), dtype=tdt, device=dev)
        d_dev = torch.empty((n_obs,), dtype=tdt, device=dev)
        oc_dev = torch.empty((n_obs, n_coord), dtype=f64, device=dev)
    if world > 1:
        dist.broadcast(gc_dev, 0)          # the grid is static: part of the plan, outside the timed region

    if rank == 0:
        metric = make_metric(w, data)
        mdesc = torch.tensor([metric.metric_id, metric.n_coord] + [0], dtype=f64, device=dev)
        mpar = torch.tensor(list(metric.params) + [0.0], dtype=f64, device=dev)[:1]
    else:
        mdesc = torch.zeros(3, dtype=f64, device=dev); mpar = torch.zeros(1, dtype=f64, device=dev)
    if world > 1:
        dist.broadcast(mdesc, 0); dist.broadcast(mpar, 0)
        from pytassim_b200.localization import metrics as mm
        if rank != 0:
            metric = mm.HaversineDistance(float(mpar[0])) if int(mdesc[0]) == 3 else mm.PeriodicDistance1D(float(mpar[0]))

    eng = LETKFEngine(k, 1, metric, w["radius"], inf_factor=w["rho"], dtype=tdt)
    eng.set_grid(gc_dev)
    eng.enable_timing(True)
    from pytassim_b200.parallel import ShardedAnalysis
    sharded = ShardedAnalysis(eng)
    nb = eng.n_blocks
    b0, b1 = sharded.ranges[rank]
    xa_dev = torch.empty_like(x_dev)

    kernel_ms, gram_ms, solve_ms = [], [], []

    def step(record=False):
        # rank 0 owns the inputs: broadcast obs-space arrays + state once per step (no-op for one GPU)
        sharded.broadcast_inputs([oc_dev, y_dev, d_dev, x_dev])
        eng.bin_obs(oc_dev, y_dev, d_dev)
        sharded.run(x_dev, xa_dev)
        if record:
            kernel_ms.append(eng.last_kernel_ms())
            gm, sm = eng.last_phase_ms()
            gram_ms.append(gm); solve_ms.append(sm)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # algorithmic work: sum over grid points of the local observation counts (exact, from the neighbour count kernel)
    eng.bin_obs(oc_dev, y_dev, d_dev)
    import ctypes
    from pytassim_b200 import _cabi
    counts = torch.zeros(n_grid, dtype=torch.int64, device=dev)
    namb = torch.zeros(1, dtype=torch.int64, device=dev)
    _cabi.check(eng.lib.b200da_neighbour_count(eng._plan, ctypes.c_void_p(counts.data_ptr()), ctypes.c_void_p(namb.data_ptr()),
                                               ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    order = torch.empty(n_grid, dtype=torch.int32, device=dev)
    _cabi.check(eng.lib.b200da_grid_order(eng._plan, ctypes.c_void_p(order.data_ptr()),
                                          ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    s0, s1 = eng.block_offset(b0), eng.block_offset(b1)
    my_pairs = int(counts[order[s0:s1].long()].sum().item())
    my_points = s1 - s0
    flops_gram = 2.0 * k * k * my_pairs + 2.0 * k * my_pairs
    flops_solve = (13.0 * k ** 3 + 2.0 * k * k) * my_points
    flops_local = flops_gram + flops_solve
    p_mean = float(counts.double().mean().item())
    del counts, order

    peak_burst = peak_sus = None
    if rank == 0:
        peak_burst, peak_sus = measure_fp64_peak(torch)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(args.steps):
        step(record=True)
    e1.record()
    barrier()
    elapsed_ms = e0.elapsed_time(e1)
    launches = launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([elapsed_ms, float(np.mean(kernel_ms)), float(np.mean(gram_ms)), float(np.mean(solve_ms))], dtype=f64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms, kern_ms, g_ms, s_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    ms_per_step = elapsed_ms / args.steps
    value = n_grid / (ms_per_step * 1e-3)

    # end to end through the host-buffer entry point (N = 1: rank 0 owns everything)
    e2e = None
    if world == 1:
        out_host = torch.empty_like(x_host).pin_memory()
        eng.analyse_host(x_host, oc_host, y_host, d_host, out=out_host)          # warm (allocates staging)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            eng.analyse_host(x_host, oc_host, y_host, d_host, out=out_host)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        h2d = x_host.numel() * esz + y_host.numel() * esz + d_host.numel() * esz + oc_host.numel() * 8
        e2e = {"value": n_grid / dt, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(out_host.numel() * esz),
               "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "api": "LETKFEngine.analyse_host -> b200da_letkf_host"}
        # sanity: the e2e result equals the device-resident result
        e2e["max_abs_diff_vs_device_path"] = float((out_host - xa_dev.cpu()).abs().max())
    else:
        # N > 1: inputs start in rank 0's pinned host memory, result is read back on rank 0
        def e2e_step():
            if rank == 0:
                x_dev.copy_(x_host, non_blocking=True); y_dev.copy_(y_host, non_blocking=True)
                d_dev.copy_(d_host, non_blocking=True); oc_dev.copy_(oc_host.t(), non_blocking=True)
            step()
            if rank == 0:
                out_host.copy_(xa_dev, non_blocking=True)
        out_host = torch.empty((1, k, n_grid), dtype=tdt).pin_memory() if rank == 0 else None
        e2e_step(); barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], dtype=f64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt[0])
        if rank == 0:
            h2d = x_host.numel() * esz + y_host.numel() * esz + d_host.numel() * esz + oc_host.numel() * 8
            e2e = {"value": n_grid / dt, "unit": unit, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(n_grid * k * esz),
                   "ms_per_step": dt * 1e3, "steps": args.e2e_steps, "api": "pinned host -> rank 0 -> NCCL broadcast -> analyse -> all-gather -> host"}

    if rank == 0:
        achieved = flops_gram / (g_ms * 1e-3) * 1e-12            # dominant kernel: the DMMA Gram kernel
        kt = k // 8 if k % 8 == 0 else (k + 1 + 7) // 8        # multiples of 8: the innovation row is accumulated by FMAs
        exec_ratio = ((kt * (kt + 1) // 2) * 128.0 + (2.0 * k if k % 8 == 0 else 0.0)) / (2.0 * k * k + 2.0 * k)
        tc_path = "tcgen05" in eng.kernel_name
        peak_used, peak_src = peak_sus, ("measured live: cuBLAS DGEMM 8192^3 via torch.matmul, sustained {0:.2f} / burst {1:.2f} "
                                         "TFLOP/s (FP64 DMMA pipe; MEASURED_PEAKS.json has no FP64 entry)".format(peak_sus, peak_burst))
        exec_note = ("executed DMMA FLOPs (lower-triangle 8x8 tiles of the padded [Yn; d] Gram) / algorithmic "
                     "FLOPs = {0:.3f}".format(exec_ratio))
        if tc_path:
            # FP32 plan: bf16 hi/lo split operands on the tcgen05 tensor cores -> the bf16 dense peak of MEASURED_PEAKS.json
            try:
                mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                peak_used = float(mp.get("bf16_tflops_sustained") or mp.get("bf16_tflops"))
                peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16 8192^3, seconds-long loop), of measured"
            except Exception:
                peak_used, peak_src = 1400.0, "fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md), of fallback"
            n_cols = (k + 1) * (k + 2) // 2
            n_chunks = -(-n_cols // 512)
            nc = -(-(-(-n_cols // n_chunks)) // 32) * 32
            exec_ratio = 3.0 * 2.0 * nc * n_chunks / (2.0 * k * k + 2.0 * k)
            exec_note = ("executed tensor FLOPs per accepted (grid point, obs) pair = 3 bf16 MMAs (hi*hi + hi*lo + lo*hi) x 2 x {0} "
                         "pair columns; / algorithmic FLOPs = {1:.3f} (candidates rejected per grid point but kept for the "
                         "128-point block add to the executed side)".format(nc * n_chunks, exec_ratio))
        path_tflops = flops_local / (kern_ms * 1e-3) * 1e-12      # Gram + solve kernels together
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic_{0}{1}.json".format(args.workload, "" if args.dtype == "f64" else "_f32"))
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": metric_name, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": w["desc"] if args.dtype == "f64" else w["desc"].replace("FP64", "FP32 arrays in HBM"), "n_grid": n_grid, "n_obs": n_obs, "ens_size": k, "mean_local_obs": p_mean,
                       "sharding": "grid-point blocks split contiguously over {0} GPU(s); obs broadcast from rank 0, "
                                   "analysis all-gathered".format(world),
                       "l2": "inputs exceed L2 (staged obs copy {0:.0f} MB + state {1:.0f} MB vs 126 MB)".format(
                           n_obs * eng.k * esz * 1.12 / 1e6, n_grid * k * esz / 1e6) if n_obs * k * esz > 2e8 else
                             "inputs fit in L2; no flush (workload is latency/compute bound, not DRAM bound)",
                       "kernel": eng.kernel_name},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_used, "unit": "TFLOP/s",
                         "frac": achieved / peak_used if peak_used else None, "traffic": traffic,
                         "kernel": eng.kernel_name, "kernel_ms": g_ms,
                         "algorithmic_flops_per_launch": flops_gram,
                         "flop_model": "Gram kernel: sum_g 2k^2 p_g + 2k p_g (SURVEY.md 8d; full k x k Gram counted, the kernel "
                                       "computes the lower triangle), p_g from the neighbour-count kernel, rank 0's share",
                         "fp64_dgemm_peak_tflops": peak_sus,
                         "peak_source": peak_src,
                         "kernel_share_of_step": g_ms / ms_per_step,
                         "executed_frac": (achieved * exec_ratio / peak_used) if peak_used else None,
                         "executed_note": exec_note,
                         "solve_kernel": {"name": "k_letkf_solve_ns (Newton-Schulz inverse square root + transform + update)", "kernel_ms": s_ms,
                                          "algorithmic_flops_per_launch": flops_solve,
                                          "achieved_tflops": flops_solve / (s_ms * 1e-3) * 1e-12 if s_ms > 0 else None,
                                          "share_of_step": s_ms / ms_per_step},
                         "path": {"achieved_tflops": path_tflops, "frac": path_tflops / peak_used if peak_used else None,
                                  "kernel_ms": kern_ms,
                                  "flop_model": "sum_g 2k^2 p_g + 2k p_g + 13k^3 + 2k^2 n_s (SURVEY.md 8d)"}},
            "solver": "FP64 Newton-Schulz (k x k solve and update in FP64 for both plan dtypes)",
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if cpu_info is not None:
            line["cpu_baseline"] = cpu_info
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
