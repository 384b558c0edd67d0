"""Diagnostic: accuracy of the global ETKF weights at cfg4 scale (k=100, M=1e6) against a numpy eigh solve of the same Gram."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-assimilate_b200"))
import torch
from pytassim_b200.engine import LETKFEngine
from pytassim_b200.localization.metrics import AbsDistance1D

k, n, m = 100, 2_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g = torch.Generator(device="cuda"); g.manual_seed(42)
x = torch.randn((1, k, n), dtype=torch.float64, device="cuda", generator=g)
stride = max(1, n // m)
hx = x[0, :, ::stride][:, :m]
yn = (hx - hx.mean(dim=0, keepdim=True)).contiguous()
d = torch.randn(m, dtype=torch.float64, device="cuda", generator=g) * 0.5
rho = 1.1
eng = LETKFEngine(k, 1, AbsDistance1D(), 1.0, inf_factor=rho)
gram = eng.etkf_gram(yn, d).cpu().numpy()
C = np.tril(gram[:k, :k]) + np.tril(gram[:k, :k], -1).T
b = gram[k, :k]
ev, U = np.linalg.eigh(C)
ev = np.clip(ev, 0, None) + (k - 1) / rho
wref = ((U / ev) @ U.T @ b)[:, None] + (U * np.sqrt((k - 1) / ev)) @ U.T
print("eig range", ev.min(), ev.max(), "kappa", ev.max() / ev.min())
w_ns = eng.etkf_weights(yn, d).cpu().numpy()
eng.set_solver("jacobi")
w_j = eng.etkf_weights(yn, d).cpu().numpy()
eng.set_solver("newton")
parts = [eng.etkf_gram(yn, d, obs_range=(a, min(a + m // 4 // 16 * 16, m) if i < 3 else m)) for i, a in enumerate(range(0, m // 4 // 16 * 16 * 4, m // 4 // 16 * 16))][:4]
gsum = sum(parts)
print("gram sharded vs whole rel", float(np.abs(gsum.cpu().numpy() - gram).max() / np.abs(gram).max()))
w_sh = eng.etkf_weights_from_gram(gsum, m).cpu().numpy()
w_fg = eng.etkf_weights_from_gram(torch.as_tensor(gram).cuda(), m).cpu().numpy()
sc = np.abs(wref).max()
for name, w in (("newton", w_ns), ("jacobi", w_j), ("from_gram(whole)", w_fg), ("from_gram(sharded sum)", w_sh)):
    print("{0:26s} max|W - Wref| / max|Wref| = {1:.3e}".format(name, np.abs(w - wref).max() / sc))
print("wmean scale", np.abs((U / ev) @ U.T @ b).max(), "wperts scale", sc)
