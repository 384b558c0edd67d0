#!/usr/bin/env python
"""BASELINE cfg5 — ensemble-size sweep k in {16, 32, 64, 128} x {FP32, FP64} on the cfg3 geometry (1000 x 1000 lat-lon grid,
2.5M observations, haversine GC c = 1000 km).  One JSON line per configuration: device time of the Gram and solve kernels
(CUDA events inside the library), grid points / s and the algorithmic Gram TFLOP/s.

    python tools/sweep_cfg5.py [--nlat 1000 --nlon 1000 --n-obs 2500000] [--ks 16,32,64,128] [--dtypes f32,f64] [--max-seconds 60]

FP64 at large k takes tens of seconds per analysis: with --fraction F only the first F of the grid-point blocks is analysed
and the rate is reported for that share (work per grid point is uniform on this geometry)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlat", type=int, default=1000)
    ap.add_argument("--nlon", type=int, default=1000)
    ap.add_argument("--n-obs", type=int, default=2_500_000)
    ap.add_argument("--ks", default="16,32,64,128")
    ap.add_argument("--dtypes", default="f32,f64")
    ap.add_argument("--fraction", type=float, default=1.0)
    ap.add_argument("--fraction-f64-large", type=float, default=0.1, help="share of the blocks for FP64 with k >= 64")
    args = ap.parse_args()
    import numpy as np
    import torch
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200.localization import metrics
    from pytassim_b200.testing import synthetic as syn
    p_mean = None
    for k in [int(v) for v in args.ks.split(",")]:
        data = syn.sphere_latlon(args.nlat, args.nlon, k, args.n_obs, seed=42)
        n_grid = data["state"].shape[-1]
        for dt in args.dtypes.split(","):
            tdt = torch.float64 if dt == "f64" else torch.float32
            eng = LETKFEngine(k, 1, metrics.HaversineDistance(6371.0), 1000.0, inf_factor=1.1, dtype=tdt)
            eng.set_grid(data["grid_rows"][:, 1:])
            eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
            eng.enable_timing(True)
            x = torch.as_tensor(np.ascontiguousarray(data["state"].reshape(1, k, n_grid)), dtype=tdt).cuda()
            xa = torch.zeros_like(x)
            frac = args.fraction_f64_large if (dt == "f64" and k >= 64) else args.fraction
            b1 = max(1, int(eng.n_blocks * frac))
            npts = eng.block_offset(b1)
            if p_mean is None:
                off, _, _, _, _ = LETKFEngine.neighbour_lists(eng, with_weights=False) if n_grid <= 200_000 else (None,) * 5
                p_mean = float(off[-1].item()) / n_grid if off is not None else args.n_obs * 0.022634
            for rep in range(2):
                eng.analyse(x, out=xa, blocks=(0, b1))
                ms = eng.last_kernel_ms()
                gm, sm = eng.last_phase_ms()
            flops = (2.0 * k * k + 2.0 * k) * p_mean * npts
            print(json.dumps({"workload": "cfg5", "k": k, "dtype": dt, "n_grid": n_grid, "n_obs": args.n_obs, "points_analysed": npts,
                              "kernel": eng.kernel_name, "total_ms": ms, "gram_ms": gm, "solve_ms": sm,
                              "gridpoints_per_s": npts / (ms * 1e-3), "mean_local_obs": p_mean,
                              "gram_algorithmic_tflops": flops / (gm * 1e-3) * 1e-12}), flush=True)
            del eng, x, xa
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
