"""FP32 plans of the CUDA path against the FP64 oracle / golden vectors.  Tolerance of the north star: analysis
within 1e-4 relative (here: max |diff| <= 1e-4 * max |reference|, the FP32 inputs being the FP64 inputs rounded).
The neighbour search always runs on FP64 coordinates, so the index lists stay bit-exact."""
import numpy as np
import pytest
import torch

import letkf_oracle as orc
from pytassim_b200.testing import synthetic as syn

pytestmark = pytest.mark.gpu

REL = 1e-4


def _metrics():
    from pytassim_b200.localization import metrics
    return metrics


def _run32(data, metric, radius, rho, weights=False, taper="gc"):
    from pytassim_b200.engine import LETKFEngine
    st = data["state"]
    n_slices, k = st.shape[0] * st.shape[1], st.shape[2]
    eng = LETKFEngine(k, n_slices, metric, radius, inf_factor=rho, taper=taper, dtype=torch.float32)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
    x = torch.as_tensor(st.reshape(n_slices, k, -1), dtype=torch.float32).cuda()
    out = eng.analyse(x, return_weights=weights)
    torch.cuda.synchronize()
    if weights:
        assert out[0].dtype == torch.float32 and out[1].dtype == torch.float32
        return eng, out[0].cpu().numpy().reshape(st.shape), out[1].cpu().numpy()
    assert out.dtype == torch.float32
    return eng, out.cpu().numpy().reshape(st.shape), None


def _close(got, want):
    scale = np.abs(want).max()
    err = np.abs(got.astype(np.float64) - want).max()
    assert err <= REL * scale, (err, scale)


SYNTH = {
    "cfg1_l96_n40_k50.npz": lambda g, m: (syn.lorenz96_1d(40, 50, 1, seed=42), m.PeriodicDistance1D(40.0)),
    "cfg2_l96_n2000_k40.npz": lambda g, m: (syn.lorenz96_1d(2000, 40, 2, seed=43), m.PeriodicDistance1D(2000.0)),
    "cfg3_sphere_small.npz": lambda g, m: (syn.sphere_latlon(24, 48, 50, 3000, seed=44), m.HaversineDistance(6371.0)),
    "euclid2d_k16.npz": lambda g, m: ({k: g[k] for k in ("state", "normed_perts", "normed_obs", "grid_rows", "obs_rows")},
                                      m.EuclideanDistance(2)),
}


@pytest.mark.parametrize("name", sorted(SYNTH))
def test_fp32_synthetic_configs_against_reference(golden, name):
    g = golden(name)
    data, metric = SYNTH[name](g, _metrics())
    sel = g["sel"]
    eng, xa, w = _run32(data, metric, float(g["radius"]), float(g["rho"]), weights=True)
    _close(xa[..., sel], g["analysis"])
    nw = g["weights"].shape[0]
    _close(w[sel[:nw]], g["weights"])
    # index lists do not depend on the plan dtype
    off, idx, _, amb, namb = eng.neighbour_lists()
    off, idx = off.cpu().numpy(), idx.cpu().numpy()
    for n, gi in enumerate(sel):
        np.testing.assert_array_equal(idx[off[gi]:off[gi + 1]], g["csr_idx"][g["csr_off"][n]:g["csr_off"][n + 1]])


@pytest.mark.parametrize("k,n_grid,stride,radius", [(3, 64, 1, 2.5), (16, 200, 1, 9.0), (32, 257, 1, 30.0), (40, 300, 2, 20.0),
                                                    (56, 200, 1, 10.0), (64, 300, 1, 12.0), (100, 96, 1, 6.0)])
def test_fp32_ensemble_sizes_against_oracle(k, n_grid, stride, radius):
    m = _metrics()
    data = syn.lorenz96_1d(n_grid, k, stride, seed=100 + k)
    _, xa, w = _run32(data, m.PeriodicDistance1D(float(n_grid)), radius, 1.07, weights=True)
    ref, wref = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                   data["obs_rows"], orc.make_dist_periodic1d(float(n_grid)), radius, inf_factor=1.07)
    _close(xa, ref)
    _close(w, wref)


def test_fp32_sphere_many_local_obs():
    """Thousands of local observations per grid point (the regime of BASELINE cfg3/cfg5) in FP32."""
    m = _metrics()
    data = syn.sphere_latlon(40, 80, 32, 40_000, seed=21)
    _, xa, _ = _run32(data, m.HaversineDistance(6371.0), 1000.0, 1.1)
    sel = np.arange(0, 3200, 257)
    ref, _ = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                data["obs_rows"], orc.make_dist_haversine(6371.0), 1000.0, inf_factor=1.1, grid_subset=sel)
    _close(xa[..., sel], ref)


def test_fp32_global_etkf_and_host_path(golden):
    g = golden("fixture_letkf.npz")
    m = _metrics()
    from pytassim_b200.engine import LETKFEngine
    st = g["state"][:, 2:]
    eng = LETKFEngine(10, 2, m.AbsDistance1D(), 1.0, inf_factor=1.0, dtype=torch.float32)
    w = eng.etkf_weights(g["b_perts"], g["b_innov"])
    assert w.dtype == torch.float32
    _close(w.cpu().numpy(), g["b_weights"])
    xa = eng.apply_weights(torch.as_tensor(st.reshape(2, 10, 40), dtype=torch.float32).cuda(), w).cpu().numpy()
    _close(xa.reshape(st.shape), g["b_analysis"])
    # host-buffer entry point in FP32 == device path
    data = syn.sphere_latlon(12, 24, 20, 800, seed=9, n_slices=2)
    s2 = data["state"]
    eng2, xa2, _ = _run32(data, m.HaversineDistance(6371.0), 1500.0, 1.1)
    out = eng2.analyse_host(s2.reshape(2, 20, -1).astype(np.float32), data["obs_rows"][:, 1:],
                            data["normed_perts"].astype(np.float32), data["normed_obs"].astype(np.float32))
    assert out.dtype == np.float32
    np.testing.assert_array_equal(out.reshape(s2.shape), xa2)


def test_fp32_edge_cases_no_obs_tiny_grid_slices():
    """tcgen05 path corner cases: no observations at all, a grid smaller than one 128-point block, the smallest
    tensor-core ensemble size (k = 8)."""
    m = _metrics()
    from pytassim_b200.engine import LETKFEngine
    data = syn.lorenz96_1d(37, 8, 1, seed=31)
    st = data["state"]
    n_slices, k = st.shape[0] * st.shape[1], st.shape[2]
    eng = LETKFEngine(k, n_slices, m.PeriodicDistance1D(37.0), 4.0, inf_factor=1.21, dtype=torch.float32)
    assert "tcgen05" in eng.kernel_name
    eng.set_grid(data["grid_rows"][:, 1:])
    x = torch.as_tensor(st.reshape(n_slices, k, -1), dtype=torch.float32).cuda()
    # (i) no observations: inflated prior  (core/etkf.py:91-95)
    eng.bin_obs(np.zeros((0, 1)), np.zeros((k, 0)), np.zeros((0,)))
    xa0 = eng.analyse(x).cpu().numpy().reshape(st.shape)
    mean = st.mean(axis=2, keepdims=True)
    _close(xa0, mean + 1.1 * (st - mean))
    # (ii) with observations, against the oracle
    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
    xa = eng.analyse(x).cpu().numpy().reshape(st.shape)
    ref, _ = orc.letkf_analysis(st, data["normed_perts"], data["normed_obs"], data["grid_rows"], data["obs_rows"],
                                orc.make_dist_periodic1d(37.0), 4.0, inf_factor=1.21)
    _close(xa, ref)
