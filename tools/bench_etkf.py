#!/usr/bin/env python
"""BASELINE cfg4 — global ETKF without localization: state of N grid points x k members, M observations.

    python tools/bench_etkf.py [--n-grid 10000000] [--k 100] [--n-obs 1000000] [--dtype f64|f32] [--steps 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_etkf.py ...
        (observation-sharded Gram -> all-reduce of (k+1)^2 doubles -> redundant solve -> state-sharded update -> all-gather;
         pytassim_b200.parallel.ShardedETKF)

Times `b200da_etkf_weights` (split-M DMMA Gram + one ensemble-space solve; pytassim/interface/etkf.py:99-120,
core/etkf.py:79-103) and `b200da_apply_weights` (x_a = mean + (x - mean) W; interface/base.py:257-278) with CUDA events on
device-resident synthetic data and prints one JSON line with the HBM / FP64 rooflines of the update kernel."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-grid", type=int, default=10_000_000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--n-obs", type=int, default=1_000_000)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    import torch
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200.localization.metrics import AbsDistance1D
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return main_sharded(args, world)
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    esz = 8 if args.dtype == "f64" else 4
    k, n, m = args.k, args.n_grid, args.n_obs
    g = torch.Generator(device="cuda"); g.manual_seed(42)
    x = torch.randn((1, k, n), dtype=tdt, device="cuda", generator=g)
    stride = max(1, n // m)
    hx = x[0, :, ::stride][:, :m].to(torch.float64)                   # identity H on every stride-th element
    yn = (hx - hx.mean(dim=0, keepdim=True)).to(tdt).contiguous()     # R = I
    d = (torch.randn(m, dtype=torch.float64, device="cuda", generator=g) * 0.5).to(tdt)
    del hx
    eng = LETKFEngine(k, 1, AbsDistance1D(), 1.0, inf_factor=1.1, dtype=tdt)
    xa = torch.empty_like(x)

    def step():
        w = eng.etkf_weights(yn, d)
        eng.apply_weights(x, w, out=xa)
        return w
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(True) for _ in range(3)]
    tw = ta = 0.0
    for _ in range(args.steps):
        ev[0].record(); w = eng.etkf_weights(yn, d); ev[1].record(); eng.apply_weights(x, w, out=xa); ev[2].record()
        ev[2].synchronize()
        tw += ev[0].elapsed_time(ev[1]); ta += ev[1].elapsed_time(ev[2])
    tw /= args.steps; ta /= args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    upd_bytes = 2.0 * k * n * esz
    upd_flops = 2.0 * k * k * n
    gram_bytes = (k + 1.0) * m * esz
    line = {
        "metric": "etkf_analysed_state_elements_per_sec", "value": n * k / ((tw + ta) * 1e-3), "unit": "state elements/s",
        "ms_per_step": tw + ta, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "cfg4: global ETKF, N={0} grid points x k={1} members, M={2} observations (every {3}th element, R = I), "
                               "inf_factor 1.1".format(n, k, m, stride)},
        "weights": {"ms": tw, "bytes": gram_bytes, "achieved_gbs": gram_bytes / (tw * 1e-3) * 1e-9,
                    "flops": 2.0 * (k + 1) ** 2 * m, "note": "split-M DMMA Gram of [Yn; d] + one k x k solve"},
        "update": {"ms": ta, "bytes": upd_bytes, "achieved_gbs": upd_bytes / (ta * 1e-3) * 1e-9,
                   "hbm_peak_gbs": hbm, "hbm_frac": upd_bytes / (ta * 1e-3) * 1e-9 / hbm,
                   "flops": upd_flops, "achieved_tflops": upd_flops / (ta * 1e-3) * 1e-12,
                   "note": "x_a = mean + (x - mean) W; algorithmic bytes = read + write of the state; 2 k^2 FLOP per grid point"},
    }
    print(json.dumps(line), flush=True)


def main_sharded(args, world):
    """N ranks: every rank builds the same synthetic arrays (same seed), reads only its observation / state ranges."""
    import torch
    import torch.distributed as dist
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200.localization.metrics import AbsDistance1D
    from pytassim_b200.parallel import ShardedETKF
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # NCCL writes its version banner to STDOUT when NCCL_DEBUG is VERSION: keep stdout to the one JSON line
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    k, n, m = args.k, args.n_grid, args.n_obs
    g = torch.Generator(device="cuda"); g.manual_seed(42)
    x = torch.randn((1, k, n), dtype=tdt, device="cuda", generator=g)
    stride = max(1, n // m)
    hx = x[0, :, ::stride][:, :m].to(torch.float64)
    yn = (hx - hx.mean(dim=0, keepdim=True)).to(tdt).contiguous()
    d = (torch.randn(m, dtype=torch.float64, device="cuda", generator=g) * 0.5).to(tdt)
    del hx
    eng = LETKFEngine(k, 1, AbsDistance1D(), 1.0, inf_factor=1.1, dtype=tdt)
    sh = ShardedETKF(eng)
    xa = torch.empty_like(x)
    res = {}
    for gather in (False, True):
        for _ in range(args.warmup):
            sh.run(x, yn, d, xa, gather=gather)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(args.steps):
            sh.run(x, yn, d, xa, gather=gather)
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[gather] = float(t.item())
    # every rank holds the whole analysis after the gathered run: compare with the unsharded call on this rank
    ref = eng.apply_weights(x, eng.etkf_weights(yn, d))
    err = float((xa - ref).abs().max().item() / ref.abs().max().item())
    if rank == 0:
        print(json.dumps({
            "metric": "etkf_analysed_state_elements_per_sec", "value": n * k / (res[True] * 1e-3), "unit": "state elements/s",
            "n_gpus": world, "ms_per_step": res[True], "ms_per_step_no_gather": res[False],
            "value_no_gather": n * k / (res[False] * 1e-3), "dtype": args.dtype, "data": "synthetic", "scaling": "strong",
            "max_rel_diff_vs_unsharded": err,
            "config": {"workload": "cfg4: global ETKF, N={0} grid points x k={1} members, M={2} observations, inf_factor 1.1; "
                                   "observation-sharded Gram, all-reduce, redundant solve, state-sharded update{3}".format(
                                       n, k, m, ", all-gather of the analysis")}}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
