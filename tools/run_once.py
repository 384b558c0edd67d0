#!/usr/bin/env python
"""One LETKF analysis of a bench workload on cuda:0 — the short command wrapped by ncu (profiles/README.md).

    python tools/run_once.py --workload cfg3 [--blocks 0:16000] [--repeat 1] [--solver newton|jacobi]

Prints the device time of the Gram and solve kernels (CUDA events inside the library)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="small")
    ap.add_argument("--blocks", default=None, help="b0:b1 range of grid-point blocks (default: all)")
    ap.add_argument("--repeat", type=int, default=1)
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--solver", default="newton", choices=["newton", "jacobi"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from pytassim_b200.engine import LETKFEngine
    w, data = bench.make_workload(args.workload)
    metric = bench.make_metric(w, data)
    k = w["k"]
    n_grid = data["state"].shape[-1]
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    eng = LETKFEngine(k, 1, metric, w["radius"], inf_factor=w["rho"], dtype=tdt)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
    eng.enable_timing(True)
    eng.set_solver(args.solver)
    if args.stats:
        eng.collect_stats(True)
    x = torch.as_tensor(np.ascontiguousarray(data["state"].reshape(1, k, n_grid)), dtype=tdt).cuda()
    xa = torch.zeros_like(x)
    blocks = None
    if args.blocks:
        b0, b1 = (int(v) for v in args.blocks.split(":"))
        blocks = (b0, min(b1, eng.n_blocks))
    for _ in range(args.repeat):
        eng.analyse(x, out=xa, blocks=blocks)
        ms = eng.last_kernel_ms()
        g, s = eng.last_phase_ms()
        npts = (eng.block_offset(blocks[1]) - eng.block_offset(blocks[0])) if blocks else n_grid
        print("kernel={0} blocks={1} points={2} total_ms={3:.3f} gram_ms={4:.3f} solve_ms={5:.3f}".format(
            eng.kernel_name, blocks or (0, eng.n_blocks), npts, ms, g, s), flush=True)
        if args.stats:
            print(eng.stats(), flush=True)


if __name__ == "__main__":
    main()
