#!/usr/bin/env python
"""End to end through the reference-facing API: ``LETKF(...).assimilate(state, observations)`` on the cfg3 shape
(1000 x 1000 lat-lon MultiIndex grid, k = 50, 2.5M observations with a nearest-grid-point operator, haversine Gaspari-Cohn
c = 1000 km), host objects in, host object out.  Two operator paths: the device gather (operator exposes ``device_index``,
SURVEY.md 8f-2) and the reference's host path (operator called on the host, HX uploaded).

    python tools/bench_interface.py [--nlat 1000 --nlon 1000 --k 50 --n-obs 2500000] [--dtype f64|f32] [--steps 2]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlat", type=int, default=1000)
    ap.add_argument("--nlon", type=int, default=1000)
    ap.add_argument("--k", type=int, default=50)
    ap.add_argument("--n-obs", type=int, default=2_500_000)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--paths", default="device_operator,host_operator")
    ap.add_argument("--profile", action="store_true", help="cProfile of the last call of the first path (host-side breakdown)")
    args = ap.parse_args()
    import torch
    from pytassim_b200 import xrlite
    from pytassim_b200.interface import LETKF
    from pytassim_b200.localization import GaspariCohn, HaversineDistance
    from pytassim_b200.obs_ops import PositionOperator
    from pytassim_b200.testing import synthetic as syn
    data = syn.sphere_latlon(args.nlat, args.nlon, args.k, args.n_obs, seed=42)
    n_grid, m = data["grid_rows"].shape[0], data["obs_rows"].shape[0]
    t = pd.to_datetime(["1992-12-25 08:00"])
    grid = pd.MultiIndex.from_arrays([data["grid_rows"][:, 1], data["grid_rows"][:, 2]], names=["lat", "lon"])
    ogrid = pd.MultiIndex.from_arrays([data["obs_rows"][:, 1], data["obs_rows"][:, 2]], names=["lat", "lon"])
    state = xrlite.DataArray(data["state"], dict(var_name=["x"], time=t, ensemble=np.arange(args.k), grid=grid),
                             ("var_name", "time", "ensemble", "grid"))
    y = np.random.RandomState(7).normal(size=(1, m))

    def observations(device):
        ds = xrlite.Dataset({
            "observations": xrlite.DataArray(y, dict(time=t, obs_grid_1=ogrid), ("time", "obs_grid_1")),
            "covariance": xrlite.DataArray(np.ones(m), dict(obs_grid_1=ogrid), ("obs_grid_1",))})
        op = PositionOperator(data["h_index"])
        ds.obs.operator = op if device else (lambda o, s: op(o, s))
        return ds
    alg = LETKF(localization=GaspariCohn(1000.0, HaversineDistance(6371.0)), inf_factor=1.1)
    alg.dtype = torch.float64 if args.dtype == "f64" else torch.float32
    out = {}
    paths = [p for p in args.paths.split(",") if p]
    for name, device in (("device_operator", True), ("host_operator", False)):
        if name not in paths:
            continue
        obs = observations(device)
        times = []
        for _ in range(args.steps + 1):                       # first call builds the plan and bins the grid
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ana = alg.assimilate(state, obs)
            torch.cuda.synchronize(); times.append(time.perf_counter() - t0)
        eng = next(iter(alg._engines.values()))
        dev_ms = eng.last_kernel_ms() if hasattr(eng, "last_kernel_ms") else None
        out[name] = dict(first_call_s=times[0], s_per_call=float(np.mean(times[1:])),
                         gridpoints_per_s=n_grid / float(np.mean(times[1:])), device_kernel_ms_last_call=dev_ms)
        if args.profile and name == paths[0]:
            import cProfile, pstats, io
            pr = cProfile.Profile(); pr.enable(); alg.assimilate(state, obs); torch.cuda.synchronize(); pr.disable()
            buf = io.StringIO(); pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(35)
            sys.stderr.write(buf.getvalue())
        out[name + "_checksum"] = float(np.asarray(ana.values, dtype=np.float64).sum())
    print(json.dumps({
        "metric": "letkf_analysed_gridpoints_per_sec", "unit": "gridpoints/s", "dtype": args.dtype, "data": "synthetic",
        "api": "pytassim_b200.interface.LETKF.assimilate(state, observations): host objects in, host object out",
        "config": {"workload": "cfg3 shape through the interface classes", "n_grid": n_grid, "n_obs": m, "ens_size": args.k},
        "value": out[paths[0]]["gridpoints_per_s"], **out,
        "same_analysis": (out["device_operator_checksum"] == out["host_operator_checksum"]) if len(paths) == 2 else None}), flush=True)


if __name__ == "__main__":
    main()
