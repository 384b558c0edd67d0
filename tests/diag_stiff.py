"""Diagnostic (test infrastructure, run by hand on the GPU box: python tests/diag_stiff.py): LETKF weights on stiff
ensemble-space problems, Newton-Schulz (two-level) and Jacobi solver against the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import letkf_oracle as orc
from pytassim_b200.engine import LETKFEngine
from pytassim_b200.localization.metrics import PeriodicDistance1D
from pytassim_b200.testing import synthetic as syn

for k, scale in [(40, 1.0), (40, 10.0), (40, 30.0), (40, 100.0), (40, 1000.0), (50, 300.0), (24, 3000.0), (100, 200.0)]:
    n_grid = 96
    data = syn.lorenz96_1d(n_grid, k, 1, seed=900 + k)
    yp = data["normed_perts"] * scale; yo = data["normed_obs"] * scale
    ref, wref = orc.letkf_analysis(data["state"], yp, yo, data["grid_rows"], data["obs_rows"],
                                   orc.make_dist_periodic1d(float(n_grid)), 6.0, inf_factor=1.05)
    out = []
    for solver in ("newton", "jacobi"):
        if solver == "jacobi" and k > 111:
            continue
        eng = LETKFEngine(k, 1, PeriodicDistance1D(float(n_grid)), 6.0, inf_factor=1.05)
        eng.set_solver(solver)
        eng.set_grid(data["grid_rows"][:, 1:]); eng.bin_obs(data["obs_rows"][:, 1:], yp, yo)
        xa, w = eng.analyse(torch.as_tensor(data["state"].reshape(1, k, n_grid)).cuda(), return_weights=True)
        w = w.cpu().numpy(); xa = xa.cpu().numpy().reshape(data["state"].shape)
        out.append((solver, np.abs(w - wref).max() / max(1.0, np.abs(wref).max()), np.abs(xa - ref).max() / np.abs(ref).max()))
    # condition number at grid point 0 from the oracle weights: Wp = sqrt(k-1) A^-1/2 -> eig(A)
    g = eng.local_gram().cpu().numpy()[0]
    c = np.tril(g[:k, :k]) + np.tril(g[:k, :k], -1).T
    ev = np.linalg.eigvalsh(c); a = (k - 1) / 1.05
    print("k=%3d scale %6.0f kappa %.1e | " % (k, scale, (ev.max() + a) / a) + " | ".join("%s W %.1e xa %.1e" % o for o in out), flush=True)
