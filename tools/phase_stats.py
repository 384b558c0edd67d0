"""Phase statistics of the fused LETKF kernel on a workload: cycles in Gram vs EVD phases, Jacobi sweeps."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))
import torch
import bench
from pytassim_b200.engine import LETKFEngine

name = sys.argv[1] if len(sys.argv) > 1 else "small"
w, data = bench.make_workload(name)
metric = bench.make_metric(w, data)
k = w["k"]
eng = LETKFEngine(k, 1, metric, w["radius"], inf_factor=w["rho"])
eng.set_grid(data["grid_rows"][:, 1:])
eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
x = torch.as_tensor(data["state"].reshape(1, k, -1)).cuda()
eng.enable_timing(True)
eng.analyse(x); torch.cuda.synchronize()
eng.collect_stats(True)
eng.analyse(x); torch.cuda.synchronize()
st = eng.stats(); ms = eng.last_kernel_ms()
nb = eng.n_blocks
out = dict(workload=name, kernel=eng.kernel_name, kernel_ms=ms, blocks=nb, n_grid=data["state"].shape[-1],
           gram_cycles_per_cta=st["gram_cycles"] / nb, evd_cycles_per_cta=st["evd_cycles"] / nb,
           setup_cycles_per_cta=st["setup_cycles"] / nb, sweeps_per_evd=st["sweeps"] / max(st["evds"], 1),
           tiles_per_cta=st["tiles"] / nb, phase_ms=eng.last_phase_ms(), jacobi_prof_cycles_per_step=[c / max(1.0, st["sweeps"] / max(st["evds"], 1) * ((k + 1) // 2 * 2 - 1) * max(1, -(-data["state"].shape[-1] // (148 * 64)))) for c in st["jacobi_prof"]])
print(json.dumps(out))
