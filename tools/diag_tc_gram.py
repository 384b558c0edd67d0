#!/usr/bin/env python
"""Accuracy of the tcgen05 Gram (FP32 plans) against the FP64 DMMA Gram on the same FP32-rounded inputs, at the operating
point of a bench workload (default: full cfg3, p ~ 56 k local observations per grid point).

    python tools/diag_tc_gram.py [--workload cfg3] [--blocks 3000:3002]

Reports, over the grid points of the chosen FP32-plan blocks: relative error of the diagonal (signed mean = accumulation bias,
max), error of the off-diagonal and innovation-row entries relative to the largest diagonal entry."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3")
    ap.add_argument("--blocks", default=None)
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from pytassim_b200.engine import LETKFEngine
    w, data = bench.make_workload(args.workload)
    metric = bench.make_metric(w, data)
    k = w["k"]
    y32 = np.ascontiguousarray(data["normed_perts"], dtype=np.float32)
    d32 = np.ascontiguousarray(data["normed_obs"], dtype=np.float32)
    e32 = LETKFEngine(k, 1, metric, w["radius"], inf_factor=w["rho"], dtype=torch.float32)
    e32.set_grid(data["grid_rows"][:, 1:])
    e32.bin_obs(data["obs_rows"][:, 1:], y32, d32)
    if args.blocks:
        b0, b1 = (int(v) for v in args.blocks.split(":"))
    else:
        b0 = e32.n_blocks // 2
        b1 = b0 + 2
    order32 = e32.grid_order().cpu().numpy()
    gp = order32[e32.block_offset(b0):e32.block_offset(b1)]
    g32 = e32.local_gram(blocks=(b0, b1))[torch.as_tensor(gp, device="cuda").long()].cpu().numpy()
    name32 = e32.kernel_name
    del e32
    torch.cuda.empty_cache()
    e64 = LETKFEngine(k, 1, metric, w["radius"], inf_factor=w["rho"], dtype=torch.float64)
    e64.set_grid(data["grid_rows"][:, 1:])
    e64.bin_obs(data["obs_rows"][:, 1:], y32.astype(np.float64), d32.astype(np.float64))
    blk = e64.blocks_of_grid(gp)
    c0, c1 = int(blk.min()), int(blk.max()) + 1
    g64 = e64.local_gram(blocks=(c0, c1))[torch.as_tensor(gp, device="cuda").long()].cpu().numpy()
    tri = np.tril(np.ones((k + 1, k + 1), dtype=bool))
    diag = np.eye(k + 1, dtype=bool)
    diag[k, k] = False
    off = tri & ~np.eye(k + 1, dtype=bool)
    off[k, :] = False
    brow = np.zeros_like(tri)
    brow[k, :k] = True
    dmax = np.abs(g64[:, diag]).max(axis=1, keepdims=True)
    err = g32 - g64
    rel_diag = err[:, diag] / g64[:, diag]
    out = {
        "kernel": name32, "blocks": [b0, b1], "n_points": int(gp.size), "fp64_blocks": [c0, c1],
        "diag_mean": float(np.mean(g64[:, diag])),
        "diag_rel_err_signed_mean": float(rel_diag.mean()), "diag_rel_err_max": float(np.abs(rel_diag).max()),
        "offdiag_err_over_diagmax_max": float((np.abs(err[:, off]) / dmax).max()),
        "offdiag_err_over_diagmax_rms": float(np.sqrt(np.mean((err[:, off] / dmax) ** 2))),
        "brow_err_over_diagmax_max": float((np.abs(err[:, brow]) / dmax).max()),
        "brow_err_over_diagmax_rms": float(np.sqrt(np.mean((err[:, brow] / dmax) ** 2))),
        "dd_rel_err": float(np.abs(err[:, k, k] / g64[:, k, k]).max()),
        "offdiag_rms_over_diagmax": float(np.sqrt(np.mean((g64[:, off] / dmax) ** 2))),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
