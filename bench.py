#!/usr/bin/env python
"""bench.py — LETKF analysed grid points / s on B200 (BASELINE.json metric), with roofline, CPU baseline and parity sample.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg1|cfg4|cfg5_k16|...] [--dtype f64|f32]
                    [--impl reference] [--no-secondary] [--secondary a,b,...]

One "step" is one full analysis of the workload: observation binning + Gram + ensemble-space solve + update (+ for N > 1
the broadcast of the observation-space arrays from rank 0 and the all-gather of the analysis; + the ambiguity protocol of
include/b200da.h, one stream synchronisation).  `value` is measured with every input already resident in HBM; `e2e` goes
through the host-buffer entry point (`b200da_letkf_host`: pinned host arrays in, host array out, copies inside the timed
region).  The headline line is cfg3 FP64 (BASELINE.json configs[2], the configuration the metric is quoted on); the default
invocation appends `secondary`: one entry per other BASELINE configuration (cfg1, cfg2, cfg3 FP32 plan, cfg4 global ETKF,
cfg5 k = 16/32/64/128 x FP32/FP64), each with its own roofline, clock record, e2e and CPU baseline.

`--impl reference` times the reference's own CPU code: the unmodified leaf modules of the reference installed under
baseline/_ref (wrapper_localization(wrapper_bridge(ETKFModule)) + GaspariCohn.localize_obs per grid point, numpy update),
loaded by file path because the package itself needs xarray / dask (absent); when baseline/_ref is missing it falls back to
the numpy port (oracle/letkf_oracle.py).  All host cores, bounded samples.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))

METRIC, UNIT = "letkf_analysed_gridpoints_per_sec", "gridpoints/s"

WORKLOADS = {
    "cfg3": dict(desc="cfg3: synthetic global 1000x1000 lat-lon grid (1M points), k=50, 2.5M obs uniform on the sphere, "
                      "haversine GC c=1000 km, inf_factor 1.1, FP64",
                 kind="sphere", nlat=1000, nlon=1000, k=50, n_obs=2_500_000, radius=1000.0, rho=1.1),
    "cfg2": dict(desc="cfg2: Lorenz-96 ring N=100k, k=40, every 2nd variable observed, periodic GC c=20, inf_factor 1.1, FP64",
                 kind="ring", n_grid=100_000, k=40, stride=2, radius=20.0, rho=1.1),
    "cfg1": dict(desc="cfg1: Lorenz-96 ring N=40, k=50, all observed, periodic GC c=5, inf_factor 1.1, FP64 "
                      "(examples/benchmark_letkf.py shape)",
                 kind="ring", n_grid=40, k=50, stride=1, radius=5.0, rho=1.1),
    "small": dict(desc="small: 100x200 lat-lon grid, k=50, 50k obs, haversine GC c=1000 km (smoke-sized cfg3), FP64",
                  kind="sphere", nlat=100, nlon=200, k=50, n_obs=50_000, radius=1000.0, rho=1.1),
    "cfg4": dict(desc="cfg4: global ETKF without localization, state 10M elements x k=100 members, 1M observations (every 10th "
                      "element, R = I), inf_factor 1.1, FP64",
                 kind="global", n_grid=10_000_000, k=100, n_obs=1_000_000, rho=1.1),
    "cfg4small": dict(desc="cfg4small: global ETKF, state 200k elements x k=100, 20k observations, FP64 (smoke-sized cfg4)",
                      kind="global", n_grid=200_000, k=100, n_obs=20_000, rho=1.1),
}
for _k in (16, 32, 64, 128):
    WORKLOADS["cfg5_k{0}".format(_k)] = dict(
        desc="cfg5: cfg3 geometry (1000x1000 lat-lon grid, 2.5M obs, haversine GC c=1000 km, inf_factor 1.1) with k={0}, "
             "FP64".format(_k),
        kind="sphere", nlat=1000, nlon=1000, k=_k, n_obs=2_500_000, radius=1000.0, rho=1.1)

# default secondary entries: (key, workload, dtype, share of the grid-point blocks analysed per step)
SECONDARY = [
    ("cfg3_f32", "cfg3", "f32", 1.0), ("cfg2_f64", "cfg2", "f64", 1.0), ("cfg2_f32", "cfg2", "f32", 1.0),
    ("cfg1_f64", "cfg1", "f64", 1.0), ("cfg4_f64", "cfg4", "f64", 1.0), ("cfg4_f32", "cfg4", "f32", 1.0),
    ("cfg5_k16_f64", "cfg5_k16", "f64", 1.0), ("cfg5_k16_f32", "cfg5_k16", "f32", 1.0),
    ("cfg5_k32_f64", "cfg5_k32", "f64", 1.0), ("cfg5_k32_f32", "cfg5_k32", "f32", 1.0),
    ("cfg5_k64_f64", "cfg5_k64", "f64", 0.25), ("cfg5_k64_f32", "cfg5_k64", "f32", 1.0),
    ("cfg5_k128_f64", "cfg5_k128", "f64", 0.10), ("cfg5_k128_f32", "cfg5_k128", "f32", 0.25),
]
SECONDARY_MULTI = ("cfg3_f32", "cfg2_f64", "cfg4_f64")       # what the N > 1 runs repeat


def desc_of(w, dtype):
    return w["desc"] if dtype == "f64" else w["desc"].replace("FP64", "FP32 arrays in HBM")


def make_workload(name):
    from pytassim_b200.testing import synthetic as syn
    w = WORKLOADS[name]
    if w["kind"] == "sphere":
        data = syn.sphere_latlon(w["nlat"], w["nlon"], w["k"], w["n_obs"], seed=42)
    elif w["kind"] == "ring":
        data = syn.lorenz96_1d(w["n_grid"], w["k"], w["stride"], seed=42)
    else:
        data = None                       # cfg4: generated on the device (8 GB state); the CPU legs draw their own sample
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


# ---- CPU legs: the reference's own leaf modules (baseline/_ref) or the numpy port, on the host cores ------------------------
_G = {}
REF_ROOT = os.path.join(ROOT, "baseline", "_ref")


def load_reference():
    """The unmodified reference leaves (core/etkf.py, core/utils.py, localization/gaspari_cohn.py, interface/wrapper.py ...)
    from baseline/_ref by file path (oracle/make_golden.py: load_reference_leaves), or None when not installed."""
    if "ref" in _G:
        return _G["ref"]
    ref = None
    if os.path.exists(os.path.join(REF_ROOT, "pytassim", "core", "etkf.py")):
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import make_golden
            ref = make_golden.load_reference_leaves(REF_ROOT)
        except Exception as exc:                     # pragma: no cover - reported in the JSON line
            sys.stderr.write("[bench] reference leaves not loadable: {0!r}\n".format(exc))
            ref = None
    _G["ref"] = ref
    return ref


def _cpu_init():
    try:
        from threadpoolctl import threadpool_limits
        _G["limit"] = threadpool_limits(limits=1)       # reference advice: OMP_NUM_THREADS=1 per worker (docs etkf.rst:55-57)
        import torch
        torch.set_num_threads(1)
    except Exception:
        pass


def _cpu_chunk(sel):
    orc, dist, w, data, ref = _G["orc"], _G["dist"], _G["w"], _G["data"], _G.get("ref_use")
    if ref is not None:
        # the reference's hot loop with the reference's own functions (interface/letkf.py:127-143): per grid point
        # wrapper_localization(wrapper_bridge(ETKFModule), GaspariCohn)(grid_row, Yn, d, obs_info=...), then the numpy form of
        # _apply_weights (interface/base.py:257-278)
        import torch
        if "ref_module" not in _G or _G.get("ref_module_w") is not w:
            module = ref.core_etkf.ETKFModule(inf_factor=torch.tensor(w["rho"], dtype=torch.float64))
            bridged = ref.wrapper.wrapper_bridge(module, torch.device("cpu"), torch.float64)
            _G["ref_module"] = ref.wrapper.wrapper_localization(bridged, ref.loc_gc.GaspariCohn((w["radius"],), dist))
            _G["ref_module_w"] = w
        localized = _G["ref_module"]
        weights = np.stack([localized(data["grid_rows"][g], data["normed_perts"], data["normed_obs"], obs_info=data["obs_rows"])
                            for g in sel], axis=0)
        st = data["state"][..., sel]
        mean = st.mean(axis=2, keepdims=True)
        ana = mean + np.einsum("vtig,gij->vtjg", st - mean, weights)
    else:
        ana, _ = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                    data["obs_rows"], dist, w["radius"], inf_factor=w["rho"], grid_subset=sel)
    return float(ana.sum())


def cpu_reference_rate(w, data, target_seconds=15.0, max_points=None, use_reference=True):
    """Grid points / s of the reference loop on all host cores; bounded sample of evenly spaced grid points."""
    import multiprocessing as mp
    orc, dist = oracle_dist(w, data)
    ref = load_reference() if use_reference else None
    _G.update(orc=orc, dist=dist, w=w, data=data, ref_use=ref)
    cores = len(os.sched_getaffinity(0))
    n_grid = data["state"].shape[-1]
    probe = np.linspace(0, n_grid - 1, 2, dtype=np.int64)
    _cpu_chunk(probe[:1])
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
    what = ("the reference's own leaf modules from baseline/_ref (GaspariCohn.localize_obs over all obs -> "
            "wrapper_localization(wrapper_bridge(ETKFModule)) -> numpy update; torch.symeig shimmed to torch.linalg.eigh)"
            if ref is not None else
            "numpy+LAPACK port of the reference loop (localize_obs over all obs -> sqrt(w) gather -> ETKF weights -> update)")
    return dict(value=len(sel) * passes / dt, unit=UNIT, cores=cores, kind="reference" if ref is not None else "port",
                sample="{0} evenly spaced grid points of {1}{4} on {2} processes x 1 thread, {5}, {3:.1f} s wall".format(
                    len(sel), n_grid, cores, dt, " x {0} passes".format(passes) if passes > 1 else "", what))


def cpu_etkf_rate(w, sample_cols=1_000_000, use_reference=True):
    """cfg4 on the host: the reference's ETKFModule on the whole (k, M) observation arrays with all cores as torch threads
    (one call, interface/etkf.py:99-120), then the numpy einsum update (interface/base.py:257-278) on a sample of the state
    columns, extrapolated linearly to N."""
    import torch
    k, n, m = w["k"], w["n_grid"], w["n_obs"]
    cores = len(os.sched_getaffinity(0))
    rnd = np.random.RandomState(42)
    cols = min(n, sample_cols)
    x = rnd.normal(size=(1, 1, k, cols))
    hx = rnd.normal(size=(k, m))
    yn = hx - hx.mean(axis=0, keepdims=True)
    d = rnd.normal(size=m) * 0.5
    ref = load_reference() if use_reference else None
    torch.set_num_threads(cores)
    t0 = time.perf_counter()
    if ref is not None:
        module = ref.core_etkf.ETKFModule(inf_factor=torch.tensor(w["rho"], dtype=torch.float64))
        W = module(torch.from_numpy(yn), torch.from_numpy(d)).numpy()
    else:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import letkf_oracle as orc
        W = orc.etkf_weights(yn, d, w["rho"])
    t_w = time.perf_counter() - t0
    t0 = time.perf_counter()
    mean = x.mean(axis=2, keepdims=True)
    ana = mean + np.einsum("vtig,ij->vtjg", x - mean, W)
    t_u = (time.perf_counter() - t0) * (n / cols)
    assert np.isfinite(ana).all()
    return dict(value=n / (t_w + t_u), unit=UNIT, cores=cores, kind="reference" if ref is not None else "port",
                sample="ETKF weights from the full (k={0}, M={1}) observation arrays ({2:.2f} s, torch with {3} threads) + numpy "
                       "einsum update on {4} of {5} state columns extrapolated linearly ({6:.2f} s for all)".format(
                           k, m, t_w, cores, cols, n, t_u))


# ---- clocks ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, period_ms=int(os.environ.get("B200DA_CLOCK_PERIOD_MS", "100"))):
        self.rows, self.proc, self.gpu, self.period = [], None, gpu_index, period_ms

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:      # the first sample marks the sampler as running
                time.sleep(0.01)
            self.first = len(self.rows)
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"], samples=0)
        self.proc.terminate()
        rows = self.rows[max(0, getattr(self, "first", 1) - 1):]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        smax = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(smax) if smax else None,
                    reasons=sorted(reasons), samples=len(sm), power_w_max=max(pw) if pw else None)


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


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


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


def _log(msg):
    sys.stderr.write("[bench] " + msg + "\n"); sys.stderr.flush()


class Ctx(object):
    """Process-wide state of the B200 arm: rank / world, device, measured peaks."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.peak_burst = self.peak_sus = None

    def init_cuda(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            # NCCL writes its version banner to STDOUT when NCCL_DEBUG is VERSION: keep stdout to the one JSON line
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=self.dev)
        if self.rank == 0:
            self.peak_burst, self.peak_sus = measure_fp64_peak(torch)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]


def sample_blocks(n_blocks, fraction, pieces=8):
    """Evenly spaced block ranges that together cover ``fraction`` of the blocks (work per block is uniform on the sphere
    only in the mean: spread the sample from pole to pole)."""
    if fraction >= 1.0:
        return [(0, n_blocks)]
    per = max(1, int(n_blocks * fraction / pieces))
    starts = np.linspace(0, n_blocks - per, pieces, dtype=np.int64)
    return [(int(s), int(s) + per) for s in starts]


# ---- one localized workload on the B200 ------------------------------------------------------------------------------
def run_letkf(ctx, wname, dtype, steps, warmup, e2e_steps, data=None, fraction=1.0, parity_points=8, cpu_info=None,
              min_seconds=0.0):
    torch, dist = ctx.torch, ctx.dist
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    from pytassim_b200.engine import LETKFEngine, launch_count
    from pytassim_b200.parallel import ShardedAnalysis
    from pytassim_b200.localization import metrics as mm
    w = WORKLOADS[wname]
    if rank == 0 and data is None:
        w, data = make_workload(wname)
    k = w["k"]
    f64 = torch.float64
    tdt = torch.float64 if dtype == "f64" else torch.float32
    ndt = np.float64 if dtype == "f64" else np.float32
    esz = 8 if dtype == "f64" else 4
    if world > 1:
        shp = torch.zeros(4, dtype=torch.int64, device=dev)
        if rank == 0:
            shp[:] = torch.tensor([data["state"].shape[-1], data["normed_obs"].shape[0], data["grid_rows"].shape[1] - 1, 0])
        dist.broadcast(shp, 0)
        n_grid, n_obs, n_coord = int(shp[0]), int(shp[1]), int(shp[2])
    else:
        n_grid, n_obs, n_coord = data["state"].shape[-1], data["normed_obs"].shape[0], data["grid_rows"].shape[1] - 1

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
    inbuf = None
    if world > 1:
        # the per-analysis inputs are views of ONE flat buffer: they reach the other ranks with one transport call per step
        from pytassim_b200.parallel import InputBuffer
        inbuf = InputBuffer([((n_obs, n_coord), f64), ((k, n_obs), tdt), ((n_obs,), tdt), ((1, k, n_grid), tdt)], dev, world,
                            rank=rank, mode=os.environ.get("B200DA_INPUT_TRANSPORT", "broadcast"))
        if rank == 0:
            for v, src in zip(inbuf.views, (oc_dev, y_dev, d_dev, x_dev)):
                v.copy_(src)
        oc_dev, y_dev, d_dev, x_dev = inbuf.views
        dist.broadcast(gc_dev, 0)          # the grid is static: part of the plan, outside the timed region
    period = float(n_grid)
    metric = mm.HaversineDistance(6371.0) if w["kind"] == "sphere" else mm.PeriodicDistance1D(period)

    eng = LETKFEngine(k, 1, metric, w["radius"], inf_factor=w["rho"], dtype=tdt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.set_grid(gc_dev)
    torch.cuda.synchronize()
    set_grid_ms = (time.perf_counter() - t0) * 1e3
    eng.enable_timing(True)

    # algorithmic work: local observation count of every grid point (exact, from the neighbour-count kernel)
    if world > 1:
        inbuf.broadcast(0)
    eng.bin_obs(oc_dev, y_dev, d_dev)
    counts, _ = eng.neighbour_counts()
    sharded = ShardedAnalysis(eng, weights=counts if fraction >= 1.0 else None)
    order = eng.grid_order()
    counts_sorted = counts[order.long()]

    def work_of_rank():
        """This rank's block ranges and their algorithmic work (the ranges move when the split is re-cut after a warm-up step)."""
        b0, b1 = sharded.ranges[rank]
        ranges = [(b0, b1)] if fraction >= 1.0 else sample_blocks(eng.n_blocks, fraction)
        pairs = points = 0
        for (c0, c1) in ranges:
            s0, s1 = eng.block_offset(c0), eng.block_offset(c1)
            pairs += int(counts_sorted[s0:s1].sum().item())
            points += s1 - s0
        return ranges, 2.0 * k * k * pairs + 2.0 * k * pairs, (13.0 * k ** 3 + 2.0 * k * k) * points, points
    my_ranges, flops_gram, flops_solve, my_points = work_of_rank()
    p_mean = float(counts.double().mean().item())
    xa_dev = torch.empty_like(x_dev)
    kernel_ms, gram_ms, solve_ms, amb = [], [], [], []

    phases = os.environ.get("B200DA_BENCH_PHASES", "0") == "1"
    phase_ms = []

    def step(record=False, transport=True):
        if fraction >= 1.0:
            # rank 0 owns the inputs: broadcast obs-space arrays + state once per step (no-op for one GPU)
            if phases and record:
                pe = [torch.cuda.Event(True) for _ in range(4)]
                pe[0].record()
            if transport:
                sharded.broadcast_inputs(inbuf if inbuf is not None else [oc_dev, y_dev, d_dev, x_dev])
            if phases and record:
                pe[1].record()
            eng.bin_obs(oc_dev, y_dev, d_dev)
            if phases and record:
                pe[2].record()
            sharded.run(x_dev, xa_dev)
            if phases and record:
                pe[3].record(); pe[3].synchronize()
                phase_ms.append([pe[i].elapsed_time(pe[i + 1]) for i in range(3)] + [eng.last_kernel_ms()])
            if record:
                kernel_ms.append(eng.last_kernel_ms())
                gm, sm = eng.last_phase_ms()
                gram_ms.append(gm); solve_ms.append(sm); amb.append(dict(eng.last_ambiguous))
        else:
            eng.bin_obs(oc_dev, y_dev, d_dev)
            km = gm = sm = 0.0
            for (c0, c1) in my_ranges:
                eng.analyse(x_dev, out=xa_dev, blocks=(c0, c1))
                if record:
                    km += eng.last_kernel_ms(); a, b = eng.last_phase_ms(); gm += a; sm += b
            if record:
                kernel_ms.append(km); gram_ms.append(gm); solve_ms.append(sm); amb.append(dict(eng.last_ambiguous))

    for _ in range(max(warmup, 1)):
        step()
        if world > 1 and fraction >= 1.0 and n_grid >= 100_000:
            sharded.rebalance(eng.last_kernel_ms())      # feedback on the work split: same network and grid every step
    my_ranges, flops_gram, flops_solve, my_points = work_of_rank()
    del counts_sorted, order
    ctx.barrier()
    if min_seconds > 0.0:                       # short workloads: enough steps for the clock sampler to see the run
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); step(); e1.record(); ctx.barrier()
        est = ctx.max_over_ranks([e0.elapsed_time(e1)])[0]
        steps = int(max(steps, min(20000, math.ceil(min_seconds * 1e3 / max(est, 1e-3)))))
    sampler = ClockSampler(ctx.local_rank).start() if rank == 0 else None
    ctx.barrier()                                # rank 0 waited for the sampler's first row: nobody starts the clock early
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        step(record=True)
    e1.record()
    ctx.barrier()
    launches = launch_count() - l0
    if phases and phase_ms:
        pm = np.mean(np.asarray(phase_ms), axis=0)
        _log("rank {0} phases ms: transport {1:.2f} bin_obs {2:.2f} analyse+gather {3:.2f} (kernels {4:.2f})".format(rank, *pm))
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms, kern_ms, g_ms, s_ms = ctx.max_over_ranks([e0.elapsed_time(e1), float(np.mean(kernel_ms)), float(np.mean(gram_ms)),
                                                          float(np.mean(solve_ms))])
    ms_per_step = elapsed_ms / steps
    n_points = n_grid if fraction >= 1.0 else my_points
    value = n_points / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer entry point -------------------------------------------------------------
    e2e = None
    if fraction < 1.0:
        e2e = {"value": None, "unit": UNIT, "note": "block-sampled run: the host-buffer entry point analyses the whole grid; "
                                                    "not measured for this entry"}
    elif world == 1:
        out_host = torch.empty_like(x_host).pin_memory()
        eng.analyse_host(x_host, oc_host, y_host, d_host, out=out_host)          # warm (allocates staging)
        torch.cuda.synchronize()
        n_e2e = e2e_steps if min_seconds <= 0.0 else int(max(e2e_steps, min(2000, math.ceil(0.5 * min_seconds * 1e3 / ms_per_step))))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            eng.analyse_host(x_host, oc_host, y_host, d_host, out=out_host)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n_e2e
        h2d = x_host.numel() * esz + y_host.numel() * esz + d_host.numel() * esz + oc_host.numel() * 8
        e2e = {"value": n_grid / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(out_host.numel() * esz),
               "ms_per_step": dt * 1e3, "steps": n_e2e, "api": "LETKFEngine.analyse_host -> b200da_letkf_host"}
        e2e["max_abs_diff_vs_device_path"] = float((out_host - xa_dev.cpu()).abs().max())
    else:
        # N > 1: the inputs sit in host memory of the node (POSIX shared memory written by rank 0, page-locked by every rank):
        # every rank uploads 1 / N of the bytes over its own PCIe link, one in-place all-gather over NVLink completes the
        # buffer on every GPU; after the analysis every rank downloads 1 / N of the result into the shared output.  Falls back
        # to "rank 0 uploads everything, broadcast" when the shared segment cannot be page-locked.
        from multiprocessing import shared_memory, resource_tracker

        def shared_pinned(nbytes, tag):
            name = [None]
            if rank == 0:
                name[0] = "b200da_{0}_{1}".format(os.getpid(), tag)
                seg = shared_memory.SharedMemory(name=name[0], create=True, size=nbytes)
            dist.broadcast_object_list(name, 0)
            if rank != 0:
                seg = shared_memory.SharedMemory(name=name[0])
                try:
                    resource_tracker.unregister(seg._name, "shared_memory")      # rank 0 owns (and unlinks) the segment
                except Exception:
                    pass
            t = torch.from_numpy(np.ndarray((nbytes,), dtype=np.uint8, buffer=seg.buf))
            ok = int(torch.cuda.cudart().cudaHostRegister(t.data_ptr(), nbytes, 0)) == 0
            return seg, t, ok
        in_bytes = inbuf.nbytes
        out_bytes = n_grid * k * esz
        per_out = (out_bytes // world + 255) // 256 * 256
        seg_in, hin, ok_in = shared_pinned(in_bytes, "in")
        seg_out, hout, ok_out = shared_pinned(per_out * world, "out")
        flag = torch.tensor([1.0 if (ok_in and ok_out) else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        split_io = bool(flag.item() > 0.5)
        if rank == 0:                                   # the host copy of the flat input buffer, byte for byte
            off = 0
            for (shape, dt_), src in zip([((n_obs, n_coord), np.float64), ((k, n_obs), ndt), ((n_obs,), ndt), ((1, k, n_grid), ndt)],
                                         (data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"],
                                          data["state"].reshape(1, k, n_grid))):
                a = np.ascontiguousarray(src, dtype=dt_)
                hin[off:off + a.nbytes].numpy()[:] = a.reshape(-1).view(np.uint8)
                off = (off + a.nbytes + InputBuffer.ALIGN - 1) // InputBuffer.ALIGN * InputBuffer.ALIGN
        ctx.barrier()
        xa_bytes = xa_dev.view(-1).view(torch.uint8)
        per_in = in_bytes // world

        def e2e_step():
            if split_io:
                inbuf.slice_of(rank).copy_(hin[rank * per_in:(rank + 1) * per_in], non_blocking=True)
                inbuf.allgather(rank)
                step(transport=False)
                o0, o1 = min(rank * per_out, out_bytes), min((rank + 1) * per_out, out_bytes)
                if o1 > o0:
                    hout[o0:o1].copy_(xa_bytes[o0:o1], non_blocking=True)
            else:
                if rank == 0:
                    x_dev.copy_(x_host, non_blocking=True); y_dev.copy_(y_host, non_blocking=True)
                    d_dev.copy_(d_host, non_blocking=True); oc_dev.copy_(oc_host.t(), non_blocking=True)
                step()
                if rank == 0:
                    hout[:out_bytes].copy_(xa_bytes, non_blocking=True)
        e2e_step(); ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        ctx.barrier()
        dt = ctx.max_over_ranks([(time.perf_counter() - t0) / e2e_steps])[0]
        e2e_diff = None
        if rank == 0:
            got = hout[:out_bytes].numpy().view(ndt)
            e2e_diff = float(np.abs(got - xa_dev.cpu().numpy().reshape(-1)).max())
        ctx.barrier()
        for t_, seg in ((hin, seg_in), (hout, seg_out)):
            try:
                torch.cuda.cudart().cudaHostUnregister(t_.data_ptr())
            except Exception:
                pass
        del hin, hout
        if rank == 0:
            h2d = in_bytes
            e2e = {"value": n_grid / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(out_bytes),
                   "ms_per_step": dt * 1e3, "steps": e2e_steps, "max_abs_diff_vs_device_path": e2e_diff,
                   "api": ("host memory of the node (shared, page-locked) -> every rank uploads 1/N over its own PCIe link -> NVLink "
                           "all-gather -> analyse -> all-gather -> every rank downloads 1/N of the analysis") if split_io else
                          "pinned host -> rank 0 -> NCCL broadcast -> analyse -> all-gather -> host"}
        for seg in (seg_in, seg_out):
            try:
                if rank == 0:
                    seg.unlink()                        # the name goes away now, the mapping when the last view is dropped
            except Exception:
                pass
            try:
                seg.close()
            except Exception:
                pass

    # ---- parity sample against the oracle, outside every timed region (rank 0; the analysis is complete on every rank) ----
    parity = None
    if rank == 0 and parity_points > 0 and fraction >= 1.0:
        orc, odist = oracle_dist(w, data)
        sel = np.unique(np.linspace(0, n_grid - 1, parity_points, dtype=np.int64))
        o_state = data["state"].astype(ndt).astype(np.float64)
        o_y = data["normed_perts"].astype(ndt).astype(np.float64)
        o_d = data["normed_obs"].astype(ndt).astype(np.float64)
        ref, _, lists = orc.letkf_analysis(o_state, o_y, o_d, data["grid_rows"], data["obs_rows"], odist, w["radius"],
                                           inf_factor=w["rho"], grid_subset=sel, return_lists=True)
        got = xa_dev[..., torch.as_tensor(sel, device=dev)].double().cpu().numpy().reshape(ref.shape)
        off, idx, _, _, _ = eng.neighbour_lists(with_weights=False, subset=sel)
        off, idx = off.cpu().numpy(), idx.cpu().numpy()
        lists_equal = all(np.array_equal(idx[off[g]:off[g + 1]], lists[n]) for n, g in enumerate(sel))
        parity = {"n": int(sel.size), "max_rel": float(np.abs(got - ref).max() / np.abs(ref).max()),
                  "tolerance": 1e-10 if dtype == "f64" else 1e-4, "lists_equal": bool(lists_equal),
                  "n_ambiguous": int(amb[-1]["n"]) if amb else None, "n_ambiguous_flipped": int(amb[-1]["flipped"]) if amb else None,
                  "against": "oracle/letkf_oracle.py on the same (dtype-rounded) inputs, evenly spaced grid points"}

    if rank != 0:
        return None
    tc_path = "tcgen05" in eng.kernel_name
    kt = k // 8 if k % 8 == 0 else (k + 1 + 7) // 8        # multiples of 8: the innovation row is accumulated by FMAs
    exec_ratio = ((kt * (kt + 1) // 2) * 128.0 + (2.0 * k if k % 8 == 0 else 0.0)) / (2.0 * k * k + 2.0 * k)
    exec_note = ("executed DMMA FLOPs (lower-triangle 8x8 tiles of the padded [Yn; d] Gram) / algorithmic FLOPs = {0:.3f}".format(exec_ratio))
    kname = eng.kernel_name
    er = eng.extra_rows
    if er and not tc_path:
        ktd = (k + 1 - er) // 8
        exec_ratio = ((ktd * (ktd + 1) // 2) * 128.0) / (2.0 * k * k + 2.0 * k)
        exec_note = ("executed DMMA FLOPs ({0} lower-triangle 8x8 tiles of rows 0..{1}) / algorithmic FLOPs = {2:.3f}; the last {3} "
                     "row(s) of [Yn; d] are accumulated by DFMA".format(ktd * (ktd + 1) // 2, ktd * 8 - 1, exec_ratio, er))
    peak_used = ctx.peak_sus
    peak_src = ("measured live: cuBLAS DGEMM 8192^3 via torch.matmul, sustained {0:.2f} / burst {1:.2f} TFLOP/s (FP64 DMMA pipe; "
                "MEASURED_PEAKS.json has no FP64 entry)".format(ctx.peak_sus, ctx.peak_burst))
    if tc_path:
        mp = measured_peaks()
        if mp.get("bf16_tflops_sustained") or mp.get("bf16_tflops"):
            peak_used = float(mp.get("bf16_tflops_sustained") or mp.get("bf16_tflops"))
            peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16 8192^3, seconds-long loop), of measured"
        else:
            peak_used, peak_src = 1400.0, "fallback 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md), of fallback"
        import re
        mt = re.search(r"_n(\d+)x(\d+)", kname)                 # pair columns per CTA x column chunks, from the kernel name
        nc, n_chunks = int(mt.group(1)), int(mt.group(2))
        exec_ratio = 3.0 * 2.0 * nc * n_chunks / (2.0 * k * k + 2.0 * k)
        exec_note = ("executed tensor FLOPs per accepted (grid point, obs) pair = 3 bf16 MMAs (hi*hi + hi*lo + lo*hi) x 2 x {0} "
                     "pair columns; / algorithmic FLOPs = {1:.3f} (candidates rejected per grid point but kept for the "
                     "128-point block add to the executed side)".format(nc * n_chunks, exec_ratio))
    # dominant kernel of the step
    gram_dominant = g_ms >= s_ms
    latency_bound = n_grid < 148 * 8
    if gram_dominant:
        achieved = flops_gram / (g_ms * 1e-3) * 1e-12
        dom = dict(kernel=kname, kernel_ms=g_ms, algorithmic_flops_per_launch=flops_gram,
                   flop_model="Gram kernel: sum_g 2k^2 p_g + 2k p_g (SURVEY.md 8d; full k x k Gram counted, the kernel computes "
                              "the lower triangle), p_g from the neighbour-count kernel, rank 0's share")
    else:
        achieved = flops_solve / (s_ms * 1e-3) * 1e-12
        peak_used, peak_src = ctx.peak_sus, ("measured live: cuBLAS DGEMM 8192^3 via torch.matmul, sustained {0:.2f} TFLOP/s (the "
                                             "solve runs on the FP64 tensor pipe for both plan dtypes)".format(ctx.peak_sus))
        dom = dict(kernel="k_letkf_solve_ns (Newton-Schulz inverse square root + transform + update)", kernel_ms=s_ms,
                   algorithmic_flops_per_launch=flops_solve,
                   flop_model="solve kernel: (13 k^3 + 2 k^2 n_s) per grid point, LAPACK-equivalent (SURVEY.md 8d: syev 9k^3 + "
                              "2 rev_evd 4k^3 + update); the Newton-Schulz iteration executes more")
        exec_ratio, exec_note = None, "LAPACK-equivalent FLOPs; see solver"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_{0}{1}.json".format(wname, "" if dtype == "f64" else "_f32"))
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "latency" if latency_bound else "tensor", "achieved": achieved, "peak": peak_used, "unit": "TFLOP/s",
                "frac": None if latency_bound else (achieved / peak_used if peak_used else None), "traffic": traffic,
                "fp64_dgemm_peak_tflops": ctx.peak_sus, "peak_source": peak_src,
                "kernel_share_of_step": dom["kernel_ms"] / ms_per_step,
                "executed_frac": (achieved * exec_ratio / peak_used) if (peak_used and exec_ratio) else None,
                "executed_note": exec_note,
                "gram_kernel": {"name": kname, "kernel_ms": g_ms, "algorithmic_flops_per_launch": flops_gram,
                                "achieved_tflops": flops_gram / (g_ms * 1e-3) * 1e-12 if g_ms > 0 else None,
                                "share_of_step": g_ms / ms_per_step},
                "solve_kernel": {"name": "k_letkf_solve_ns (Newton-Schulz inverse square root + transform + update)", "kernel_ms": s_ms,
                                 "algorithmic_flops_per_launch": flops_solve,
                                 "achieved_tflops": flops_solve / (s_ms * 1e-3) * 1e-12 if s_ms > 0 else None,
                                 "share_of_step": s_ms / ms_per_step},
                "path": {"achieved_tflops": (flops_gram + flops_solve) / (kern_ms * 1e-3) * 1e-12, "kernel_ms": kern_ms,
                         "frac": (flops_gram + flops_solve) / (kern_ms * 1e-3) * 1e-12 / ctx.peak_sus if (ctx.peak_sus and not tc_path) else None,
                         "flop_model": "sum_g 2k^2 p_g + 2k p_g + 13k^3 + 2k^2 n_s (SURVEY.md 8d)"}}
    if latency_bound:
        roofline["note"] = ("{0} grid points = {0} CTAs of work on 148 SMs: less than one wave, launch-latency bound; time "
                            "reported, no roofline claim (SURVEY.md 8d)".format(n_grid))
    roofline.update(dom)
    if world == 1:
        l2 = ("inputs exceed L2 (staged obs copy {0:.0f} MB + state {1:.0f} MB vs 126 MB)".format(
                  n_obs * eng.k * esz * 1.12 / 1e6, n_grid * k * esz / 1e6) if n_obs * k * esz > 2e8 else
              "inputs fit in L2; no flush (workload is latency/compute bound, not DRAM bound)")
    else:
        l2 = "inputs exceed L2" if n_obs * k * esz > 2e8 else "inputs fit in L2; no flush"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": {"workload": desc_of(w, dtype), "n_grid": n_grid, "n_obs": n_obs, "ens_size": k, "mean_local_obs": p_mean,
                   "sharding": "grid-point blocks split over {0} GPU(s) in contiguous ranges balanced by local-observation "
                               "count and re-cut after every warm-up step from the ranks' measured times; inputs from rank 0 as one flat buffer "
                               "({1}), analysis all-gathered".format(
                                   world, "n/a" if inbuf is None else inbuf.mode),
                   "l2": l2, "kernel": kname},
        "roofline": roofline,
        "solver": "FP64 Newton-Schulz (k x k solve and update in FP64 for both plan dtypes)",
        "set_grid_ms": set_grid_ms,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "parity_sample": parity,
    }
    if fraction < 1.0:
        line["config"]["block_sample"] = ("{0:.0%} of the grid-point blocks per step ({1} grid points in {2} evenly spaced ranges "
                                          "from pole to pole); value = analysed grid points / time".format(fraction, my_points, len(my_ranges)))
    if cpu_info is not None:
        line["cpu_baseline"] = cpu_info
    del eng, sharded, x_dev, y_dev, d_dev, oc_dev, xa_dev, gc_dev
    torch.cuda.empty_cache()
    return line


# ---- cfg4: global ETKF -----------------------------------------------------------------------------------------------
def run_etkf(ctx, wname, dtype, steps, warmup, e2e_steps, cpu_info=None, min_seconds=0.0):
    torch, dist = ctx.torch, ctx.dist
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    from pytassim_b200.engine import LETKFEngine, launch_count
    from pytassim_b200.localization.metrics import AbsDistance1D
    from pytassim_b200.parallel import ShardedETKF
    w = WORKLOADS[wname]
    k, n, m = w["k"], w["n_grid"], w["n_obs"]
    tdt = torch.float64 if dtype == "f64" else torch.float32
    esz = 8 if dtype == "f64" else 4
    # every rank builds the same synthetic arrays (same seed) and reads only its observation / state ranges
    g = torch.Generator(device=dev); g.manual_seed(42)
    x = torch.randn((1, k, n), dtype=tdt, device=dev, generator=g)
    stride = max(1, n // m)
    hx = x[0, :, ::stride][:, :m].to(torch.float64)                   # identity H on every stride-th element
    yn = (hx - hx.mean(dim=0, keepdim=True)).to(tdt).contiguous()     # R = I
    d = (torch.randn(m, dtype=torch.float64, device=dev, generator=g) * 0.5).to(tdt)
    del hx
    eng = LETKFEngine(k, 1, AbsDistance1D(), 1.0, inf_factor=w["rho"], dtype=tdt)
    sh = ShardedETKF(eng)
    xa = sh.alloc_output(x.shape, tdt, dev)      # symmetric memory for N > 1: the update kernel writes into every rank's array
    fused_gather = world > 1 and getattr(sh, "_peers", None) is not None
    sh.fused_stores = os.environ.get("B200DA_ETKF_FUSED_STORES", "0") == "1"

    def step():
        sh.run(x, yn, d, xa, gather=True)
    for _ in range(max(warmup, 1)):
        step()
    ctx.barrier()
    if min_seconds > 0.0:
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); step(); e1.record(); ctx.barrier()
        est = ctx.max_over_ranks([e0.elapsed_time(e1)])[0]
        steps = int(max(steps, min(5000, math.ceil(min_seconds * 1e3 / max(est, 1e-3)))))
    # per-kernel times (one untimed pass, CUDA events on the launch stream)
    ev = [torch.cuda.Event(True) for _ in range(4)]
    m0, m1 = sh.ranges(m)[rank]
    c0, c1 = sh.ranges(n)[rank]
    ev[0].record(); gram = eng.etkf_gram(yn, d, obs_range=(m0, m1))
    if world > 1:
        dist.all_reduce(gram)
    ev[1].record(); wts = eng.etkf_weights_from_gram(gram, m); ev[2].record()
    eng.apply_weights_cols(x, wts, c0, c1, xa); ev[3].record(); ev[3].synchronize()
    t_gram, t_solve, t_upd = ctx.max_over_ranks([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])])
    ctx.barrier()
    sampler = ClockSampler(ctx.local_rank).start() if rank == 0 else None
    ctx.barrier()                                # rank 0 waited for the sampler's first row: nobody starts the clock early
    l0 = launch_count()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    ctx.barrier()
    launches = launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ctx.max_over_ranks([e0.elapsed_time(e1)])[0] / steps
    sharded_ms = None
    if world > 1:                                # the analysed columns left where they were computed (no copy to the other ranks)
        for _ in range(2):
            sh.run(x, yn, d, xa, gather=False)
        ctx.barrier()
        e0.record()
        for _ in range(steps):
            sh.run(x, yn, d, xa, gather=False)
        e1.record()
        ctx.barrier()
        sharded_ms = ctx.max_over_ranks([e0.elapsed_time(e1)])[0] / steps
        sh.run(x, yn, d, xa, gather=True)        # the complete analysis again for the checks below
        ctx.barrier()
    # parity property at full size: the sharded result equals the single-call result on this rank (N > 1), and the weights
    # satisfy the defining identities of core/etkf.py:57-77:  W_p^2 (C + a I) = (k - 1) I,  (C + a I) w_mean = b
    gram_full = eng.etkf_gram(yn, d).double()
    C = gram_full[:k, :k]; C = torch.tril(C) + torch.tril(C, -1).t(); b = gram_full[k, :k]
    W = eng.etkf_weights(yn, d).double()
    a = (k - 1) / w["rho"]
    A = C + a * torch.eye(k, dtype=torch.float64, device=dev)
    wbar = torch.linalg.solve(A, b)
    Wp = W - wbar[:, None]
    ident = float(((Wp @ Wp @ A) / (k - 1) - torch.eye(k, dtype=torch.float64, device=dev)).abs().max())
    ref = eng.apply_weights(x, eng.etkf_weights(yn, d))
    shard_err = float((xa - ref).abs().max() / ref.abs().max())
    del ref
    e2e = None
    if rank == 0:
        need = 2 * x.numel() * esz + yn.numel() * esz
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 0
        if world > 1:
            e2e = {"value": None, "unit": UNIT, "note": "N > 1: the e2e leg of cfg4 is measured at N = 1 only"}
        elif avail > 2.5 * need:
            xh = torch.empty(x.shape, dtype=tdt).pin_memory(); xh.copy_(x)
            yh = torch.empty(yn.shape, dtype=tdt).pin_memory(); yh.copy_(yn)
            dh = torch.empty(d.shape, dtype=tdt).pin_memory(); dh.copy_(d)
            oh = torch.empty(x.shape, dtype=tdt).pin_memory()
            xd, yd, dd = torch.empty_like(x), torch.empty_like(yn), torch.empty_like(d)

            def e2e_step():
                xd.copy_(xh, non_blocking=True); yd.copy_(yh, non_blocking=True); dd.copy_(dh, non_blocking=True)
                wt = eng.etkf_weights(yd, dd)
                eng.apply_weights(xd, wt, out=xa)
                oh.copy_(xa, non_blocking=True)
            e2e_step(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / e2e_steps
            e2e = {"value": n / dt, "unit": UNIT, "h2d_bytes_per_step": int(need - x.numel() * esz + d.numel() * esz),
                   "d2h_bytes_per_step": int(x.numel() * esz), "ms_per_step": dt * 1e3, "steps": e2e_steps,
                   "api": "pinned host -> LETKFEngine.etkf_weights + apply_weights (b200da_etkf_weights, b200da_apply_weights) -> host"}
            del xh, yh, dh, oh, xd, yd, dd
        else:
            e2e = {"value": None, "unit": UNIT, "note": "not measured: {0:.0f} GB of pinned host memory needed, {1:.0f} GB "
                                                        "available".format(need / 1e9, avail / 1e9)}
    if rank != 0:
        return None
    hbm = float(measured_peaks().get("hbm_gbs", 6550.0))
    upd_flops = 2.0 * k * k * (c1 - c0)
    upd_bytes = 2.0 * k * (c1 - c0) * esz
    line = {
        "metric": METRIC, "value": n / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype,
        "data": "synthetic (generated on the device, seed 42)",
        "state_elements_per_sec": n * k / (ms_per_step * 1e-3),
        "config": {"workload": desc_of(w, dtype), "n_grid": n, "n_obs": m, "ens_size": k,
                   "sharding": "observation-sharded Gram -> all-reduce of (k+1)^2 doubles -> redundant k x k solve -> state-sharded "
                               "update{1} ({0} GPU(s))".format(world, "" if world == 1 else (
                                   (" whose kernel stores every analysed tile into the analysis array of every rank over NVLink "
                                    "peer memory (b200da_apply_weights_cols_peers: update and all-gather in one kernel)" if sh.fused_stores
                                    else " in 8 column pieces, each pushed into the analysis array of every rank (symmetric memory, NVLink "
                                         "peer stores on a copy stream) while the next piece is computed") if fused_gather
                                   else " -> all-gather of the analysis (NCCL)")),
                   "l2": "inputs exceed L2 (state {0:.1f} GB)".format(n * k * esz / 1e9) if n * k * esz > 2e8 else "inputs fit in L2; no flush",
                   "kernel": "k_etkf_gram + k_letkf_solve_ns + k_apply_global"},
        "roofline": {"bound": "tensor", "kernel": "k_apply_global (update Xa = W'^T X as a streaming DMMA GEMM)", "kernel_ms": t_upd,
                     "achieved": upd_flops / (t_upd * 1e-3) * 1e-12, "peak": ctx.peak_sus, "unit": "TFLOP/s",
                     "frac": upd_flops / (t_upd * 1e-3) * 1e-12 / ctx.peak_sus if ctx.peak_sus else None, "traffic": None,
                     "algorithmic_flops_per_launch": upd_flops, "flop_model": "2 k^2 per state column of this rank (SURVEY.md 8d)",
                     "peak_source": "measured live: cuBLAS DGEMM 8192^3 via torch.matmul, sustained",
                     "hbm": {"achieved_gbs": upd_bytes / (t_upd * 1e-3) * 1e-9, "peak_gbs": hbm,
                             "frac": upd_bytes / (t_upd * 1e-3) * 1e-9 / hbm, "algorithmic_bytes_per_launch": upd_bytes,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs, of measured"},
                     "kernel_share_of_step": t_upd / ms_per_step,
                     "phases_ms": {"gram_plus_allreduce": t_gram, "solve": t_solve, "update": t_upd}},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "parity_sample": {"identity_residual_Wp2_A_over_km1_minus_I": ident, "sharded_vs_single_max_rel": shard_err,
                          "tolerance": 1e-10 if dtype == "f64" else 1e-4,
                          "against": "defining identities of core/etkf.py:57-77 on the device Gram (size-independent property; "
                                     "tests/test_gpu_parity.py checks the same kernels against the oracle at reduced size)"},
    }
    if sharded_ms is not None:
        line["sharded"] = {"value": n / (sharded_ms * 1e-3), "unit": UNIT, "ms_per_step": sharded_ms,
                           "note": "the same analysis with every rank keeping only its own column range (what a sharded consumer "
                                   "needs); `value` delivers the whole (k, N) analysis to every rank, which moves (N - N / G) k "
                                   "elements into each GPU over NVLink (900 GB/s per direction) - for cfg4 on 8 GPUs 7 GB = 7.8 ms, "
                                   "next to 9.4 ms for the whole HBM-bound update on one GPU"}
    if cpu_info is not None:
        line["cpu_baseline"] = cpu_info
    del x, xa, yn, d, eng
    torch.cuda.empty_cache()
    return line


def secondary_plan(args, world):
    if args.no_secondary:
        return []
    if args.secondary:
        keys = [s for s in args.secondary.split(",") if s]
        return [e for e in SECONDARY if e[0] in keys]
    if args.workload != "cfg3" or args.dtype != "f64":
        return []
    if world > 1:
        return [e for e in SECONDARY if e[0] in SECONDARY_MULTI]
    return list(SECONDARY)


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
    ap.add_argument("--no-secondary", action="store_true", help="only the headline workload")
    ap.add_argument("--secondary", default="", help="comma-separated secondary entries to run instead of the default set")
    ap.add_argument("--fraction", type=float, default=1.0, help="share of the grid-point blocks analysed per step (headline)")
    ap.add_argument("--parity-points", type=int, default=12)
    args = ap.parse_args()
    ctx = Ctx()
    rank, world = ctx.rank, ctx.world
    plan = secondary_plan(args, world)

    # ------------------------------------------------------------------------------------------------------------
    # reference arm: the reference's CPU code on the host cores (rank 0 only)
    # ------------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        w, data = make_workload(args.workload)
        vals, info = [], None
        t_all = time.perf_counter()
        if w["kind"] == "global":
            for _ in range(args.steps):
                info = cpu_etkf_rate(w); vals.append(info["value"])
        else:
            for _ in range(max(1, args.warmup > 0)):
                cpu_reference_rate(w, data, target_seconds=3.0)
            t_all = time.perf_counter()
            for _ in range(args.steps):
                info = cpu_reference_rate(w, data, target_seconds=12.0)
                vals.append(info["value"])
        v = float(np.mean(vals))
        info["value"] = v
        ms = 1000.0 * (time.perf_counter() - t_all) / args.steps
        sec = {}
        done = {}
        for key, wname, dtype, _ in plan:                     # the CPU code is FP64 throughout: one sample per workload
            if wname not in done:
                _log("reference sample " + wname)
                ws, ds = make_workload(wname)
                done[wname] = cpu_etkf_rate(ws) if ws["kind"] == "global" else cpu_reference_rate(ws, ds, target_seconds=4.0)
                del ds
            sec[key] = {"value": done[wname]["value"], "unit": UNIT, "cpu_baseline": done[wname],
                        "config": {"workload": desc_of(WORKLOADS[wname], "f64")}}
        line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "impl": "reference",
                "config": {"workload": desc_of(w, "f64"), "inputs": "host resident"},
                "cpu_baseline": info,
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        if sec:
            line["secondary"] = sec
        _emit(line)
        return

    # ------------------------------------------------------------------------------------------------------------
    # B200 arm.  Phase A (before CUDA is initialised: the CPU pool forks): CPU baselines on bounded samples, N = 1 only
    # ------------------------------------------------------------------------------------------------------------
    w = WORKLOADS[args.workload]
    datasets, cpu = {}, {}
    if rank == 0:
        names = [args.workload] + [e[1] for e in plan]
        for wname in dict.fromkeys(names):
            ww, dd = make_workload(wname)
            datasets[wname] = dd
            if world == 1 and not args.no_cpu_baseline:
                _log("cpu baseline sample " + wname)
                tgt = 12.0 if wname == args.workload else 3.0
                cpu[wname] = cpu_etkf_rate(ww) if ww["kind"] == "global" else cpu_reference_rate(ww, dd, target_seconds=tgt)
    ctx.init_cuda()

    _log("headline " + args.workload)
    if w["kind"] == "global":
        line = run_etkf(ctx, args.workload, args.dtype, args.steps, args.warmup, args.e2e_steps, cpu_info=cpu.get(args.workload))
    else:
        line = run_letkf(ctx, args.workload, args.dtype, args.steps, args.warmup, args.e2e_steps,
                         data=datasets.pop(args.workload, None) if args.workload not in [e[1] for e in plan] else datasets.get(args.workload),
                         fraction=args.fraction, parity_points=args.parity_points, cpu_info=cpu.get(args.workload))
    sec = {}
    remaining = [e[1] for e in plan]
    for key, wname, dtype, fraction in plan:
        _log("secondary " + key)
        ws = WORKLOADS[wname]
        t0 = time.perf_counter()
        try:
            if ws["kind"] == "global":
                ent = run_etkf(ctx, wname, dtype, 5, 3, 1, cpu_info=cpu.get(wname), min_seconds=2.0)
            else:
                big = ws["kind"] == "sphere"
                ent = run_letkf(ctx, wname, dtype, 3 if big else 5, 3, 1, data=datasets.get(wname), fraction=fraction,
                                parity_points=6 if big else 8, cpu_info=cpu.get(wname), min_seconds=2.0)
        except Exception as exc:                                # one failing entry must not take the headline with it
            ent = {"error": repr(exc)}
            if world > 1:
                raise
        remaining.remove(wname)
        if wname not in remaining:
            datasets.pop(wname, None)
        if ent is not None:
            ent["wall_s"] = time.perf_counter() - t0
            sec[key] = ent
    if rank == 0:
        if sec:
            line["secondary"] = sec
        _emit(line)
    if world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
