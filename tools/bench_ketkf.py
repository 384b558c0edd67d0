#!/usr/bin/env python
"""Cost of the widened ensemble-space problems next to the plain LETKF: LKETKF (kernelise pass; RBF kernel on the Newton-Schulz
solver, tanh kernel on the Jacobi solver) and one localized IEnKS iteration with per-grid incoming weights (k_ienks_pre +
solve + k_ienks_keep; transform and bundle variants) on the cfg2 shape (Lorenz-96 ring, N = 100 000, k = 40, every 2nd variable
observed) and on a reduced cfg3 shape (300 x 300 sphere grid, k = 50, 225 000 observations).  One JSON line per workload:
device time per analysis (CUDA events around the call, median of --steps after --warmup)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))


def timed(eng, x, xa, steps, warmup, call=None):
    import torch
    if call is None:
        call = lambda: eng.analyse(x, out=xa)
    for _ in range(warmup):
        call()
    ts = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        call()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--only", default=None, help="cfg2 | cfg3 | cfg3b: run one workload (the short command wrapped by ncu)")
    args = ap.parse_args()
    import numpy as np
    import torch
    from pytassim_b200 import kernels as K
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200.localization import metrics as m
    from pytassim_b200.testing import synthetic as syn
    torch.cuda.set_device(0)
    works = [("cfg2: L96 ring N=100000 k=40 M=50000 c=20", syn.lorenz96_1d(100000, 40, 2, seed=42), m.PeriodicDistance1D(100000.0), 20.0, 38),
             ("cfg3/11: sphere 300x300 k=50 M=225000 c=1000km", syn.sphere_latlon(300, 300, 50, 225000, seed=42), m.HaversineDistance(6371.0), 1000.0, 5100)]
    if args.only == "cfg3b":             # a third of cfg3 in each direction of the grid: p about 17 000 local observations
        works = [("cfg3b: sphere 550x550 k=50 M=750000 c=1000km", syn.sphere_latlon(550, 550, 50, 750000, seed=42),
                  m.HaversineDistance(6371.0), 1000.0, 17000)]
    works = [w for w in works if args.only is None or w[0].startswith(args.only)]
    for name, data, metric, radius, p in works:
        k, n = data["state"].shape[2], data["state"].shape[3]
        x = torch.as_tensor(np.ascontiguousarray(data["state"].reshape(1, k, n))).cuda()
        xa = torch.empty_like(x)
        out = {"workload": name, "n_grid": n, "ens_size": k}
        for label, kernel in (("letkf", None), ("lketkf_rbf", K.RBFKernel(gamma=0.5 / p)),
                              ("lketkf_tanh_jacobi", K.TanhKernel(coeff=1.0 / p, const=0.1))):
            eng = LETKFEngine(k, 1, metric, radius, inf_factor=1.1)
            if kernel is not None:
                eng.set_kernel(kernel)
            eng.set_grid(data["grid_rows"][:, 1:])
            eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
            ms = timed(eng, x, xa, args.steps, args.warmup)
            out[label] = {"ms_per_analysis": ms, "gridpoints_per_s": n / ms * 1e3, "kernel": eng.kernel_name}
            if kernel is None:                        # one IEnKS iteration from per-grid weights (the LETKF weights of this engine)
                _, w_in = eng.analyse(x, out=xa, return_weights=True)
                for label2, eps, scale in (("lienks_transform_step", None, 1.0), ("lienks_bundle_step", 1e-2, 1e-2)):
                    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"] * scale, data["normed_obs"])
                    ms = timed(eng, x, xa, args.steps, args.warmup, call=lambda: eng.ienks_step(x, w_in, tau=0.8, epsilon=eps, out=xa))
                    out[label2] = {"ms_per_analysis": ms, "gridpoints_per_s": n / ms * 1e3}
                del w_in
            del eng
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
