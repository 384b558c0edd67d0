"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden vectors produced by
the reference's own code.  Tolerances: local-observation index lists bit-exact; analysis and weights within
1e-10 (FP64, BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

import letkf_oracle as orc
from pytassim_b200.testing import synthetic as syn

pytestmark = pytest.mark.gpu

RTOL = 1e-10
ATOL = 1e-10
# taper values: both the reference (numpy pow) and the device (Horner) sum terms of magnitude ~10 that cancel
# towards the cutoff, so individual weights agree to ~1e-15 absolute, not relative
W_RTOL = 1e-12
W_ATOL = 5e-15


def _engine(k, n_slices, metric, radius, **kw):
    from pytassim_b200.engine import LETKFEngine
    return LETKFEngine(k, n_slices, metric, radius, **kw)


def _metrics():
    from pytassim_b200.localization import metrics
    return metrics


def _run(data, metric, radius, rho, taper="gc", eps=1e-5, weights=False, solver="newton"):
    st = data["state"]
    n_slices = st.shape[0] * st.shape[1]
    k = st.shape[2]
    eng = _engine(k, n_slices, metric, radius, inf_factor=rho, taper=taper, epsilon=eps)
    eng.set_solver(solver)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
    x = torch.as_tensor(st.reshape(n_slices, k, -1)).cuda()
    out = eng.analyse(x, return_weights=weights, count_ambiguous=True)
    torch.cuda.synchronize()
    if weights:
        xa, w, namb = out
        return eng, xa.cpu().numpy().reshape(st.shape), w.cpu().numpy(), namb
    xa, namb = out
    return eng, xa.cpu().numpy().reshape(st.shape), None, namb


def _check_lists(eng, off_ref, idx_ref, sel=None, w_ref=None):
    off, idx, w, amb, namb = eng.neighbour_lists()
    off, idx, w = off.cpu().numpy(), idx.cpu().numpy(), w.cpu().numpy()
    assert namb == 0 and not amb.any().item()
    if sel is None:
        np.testing.assert_array_equal(off, off_ref)
        np.testing.assert_array_equal(idx, idx_ref)
        if w_ref is not None:
            np.testing.assert_allclose(w, w_ref, rtol=W_RTOL, atol=W_ATOL)
        return
    for n, g in enumerate(sel):
        got = idx[off[g]:off[g + 1]]
        want = idx_ref[off_ref[n]:off_ref[n + 1]]
        np.testing.assert_array_equal(got, want)
        if w_ref is not None:
            np.testing.assert_allclose(w[off[g]:off[g + 1]], w_ref[off_ref[n]:off_ref[n + 1]], rtol=W_RTOL, atol=W_ATOL)


def test_fixture_letkf_gc10_against_reference(golden):
    """tests/unit_tests/interface/test_letkf.py:106-157 on the reference's netCDF fixtures."""
    g = golden("fixture_letkf.npz")
    m = _metrics()
    data = dict(state=g["state"][:, :1], normed_perts=g["a_perts"], normed_obs=g["a_innov"],
                grid_rows=g["a_grid_rows"], obs_rows=g["a_obs_rows"])
    eng, xa, w, namb = _run(data, m.AbsDistance1D(), (10.0,), 1.0, weights=True)
    np.testing.assert_allclose(xa, g["a_analysis"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(w, g["a_weights"], rtol=RTOL, atol=ATOL)
    _check_lists(eng, g["a_csr_off"], g["a_csr_idx"], w_ref=g["a_csr_w"])
    np.testing.assert_allclose(xa.sum(), -2.616127006116, atol=1e-9)


def test_fixture_letkf_gcinf_against_reference(golden):
    g = golden("fixture_letkf.npz")
    m = _metrics()
    data = dict(state=g["state"][:, 1:2], normed_perts=g["c_perts"], normed_obs=g["c_innov"],
                grid_rows=g["a_grid_rows"], obs_rows=g["a_obs_rows"])
    eng, xa, w, _ = _run(data, m.AbsDistance1D(), 8.0, 1.1, taper="gcinf", weights=True)
    np.testing.assert_allclose(xa, g["c_analysis"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(w, g["c_weights"], rtol=RTOL, atol=ATOL)
    _check_lists(eng, g["c_csr_off"], g["c_csr_idx"])


def test_fixture_global_etkf_against_reference(golden):
    """tests/unit_tests/interface/test_letkf.py:64-70 / test_etkf.py:136-167: global weights + update."""
    g = golden("fixture_letkf.npz")
    m = _metrics()
    st = g["state"][:, 2:]
    eng = _engine(10, 2, m.AbsDistance1D(), 1.0, inf_factor=1.0)
    w = eng.etkf_weights(g["b_perts"], g["b_innov"])
    np.testing.assert_allclose(w.cpu().numpy(), g["b_weights"], rtol=RTOL, atol=ATOL)
    xa = eng.apply_weights(torch.as_tensor(st.reshape(2, 10, 40)).cuda(), w).cpu().numpy().reshape(st.shape)
    np.testing.assert_allclose(xa, g["b_analysis"], rtol=RTOL, atol=ATOL)
    # a localization radius so large that every weight is ~1 must reproduce the global ETKF to ~1e-6 only
    # (GC(r) = 1 - 5/3 r^2 + ...), so the exact equivalence is checked with the all-ones path instead:
    xa_pg = eng.apply_weights(torch.as_tensor(st.reshape(2, 10, 40)).cuda(), w[None].repeat(40, 1, 1).contiguous())
    np.testing.assert_allclose(xa_pg.cpu().numpy().reshape(st.shape), g["b_analysis"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("n", range(8))
def test_core_weights_against_reference(golden, n):
    """ETKFModule.forward on random inputs, k in {3..100}: the device EVD + transform alone."""
    g = golden("core_random.npz")
    m = _metrics()
    Y, d, rho = g[f"Y{n}"], g[f"d{n}"], float(g[f"rho{n}"])
    eng = _engine(Y.shape[0], 1, m.AbsDistance1D(), 1.0, inf_factor=rho)
    w = eng.etkf_weights(Y, d).cpu().numpy()
    np.testing.assert_allclose(w, g[f"W{n}"], rtol=RTOL, atol=ATOL)


def test_core_kat_and_empty(golden):
    """tests/unit_tests/core/test_etkf.py:142-233: 2-member KAT and the empty-observation prior."""
    g = golden("core_kat.npz")
    m = _metrics()
    eng = _engine(2, 1, m.AbsDistance1D(), 1.0)
    w = eng.etkf_weights(g["normed_perts"], g["normed_obs"].reshape(-1)).cpu().numpy()
    np.testing.assert_allclose(w, g["W"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(w - g["w_mean"], g["w_perts"], atol=1e-12)
    eng = _engine(10, 1, m.AbsDistance1D(), 1.0, inf_factor=1.1)
    w = eng.etkf_weights(np.ones((10, 0)), np.ones((0,))).cpu().numpy()
    np.testing.assert_allclose(w, np.sqrt(1.1) * np.eye(10), atol=1e-14)
    with pytest.raises(ValueError):
        eng.etkf_weights(np.ones((10, 4)), np.ones((3,)))


SYNTH = {
    "cfg1_l96_n40_k50.npz": lambda g, m: (syn.lorenz96_1d(40, 50, 1, seed=42), m.PeriodicDistance1D(40.0)),
    "cfg2_l96_n2000_k40.npz": lambda g, m: (syn.lorenz96_1d(2000, 40, 2, seed=43), m.PeriodicDistance1D(2000.0)),
    "bench_default_n1000_k50.npz": lambda g, m: (syn.lorenz96_1d(1000, 50, 10, seed=45), m.AbsDistance1D()),
    "cfg3_sphere_small.npz": lambda g, m: (syn.sphere_latlon(24, 48, 50, 3000, seed=44), m.HaversineDistance(6371.0)),
    "euclid2d_k16.npz": lambda g, m: ({k: g[k] for k in ("state", "normed_perts", "normed_obs", "grid_rows", "obs_rows")},
                                      m.EuclideanDistance(2)),
}


@pytest.mark.parametrize("solver", ["newton", "jacobi"])
@pytest.mark.parametrize("name", sorted(SYNTH))
def test_synthetic_configs_against_reference(golden, name, solver):
    """BASELINE.json configurations (scaled) against outputs of the reference's own hot loop, for both ensemble-space
    solvers (tensor-core Newton-Schulz inverse square root, shared-memory Jacobi eigendecomposition)."""
    g = golden(name)
    data, metric = SYNTH[name](g, _metrics())
    sel = g["sel"]
    eng, xa, w, namb = _run(data, metric, float(g["radius"]), float(g["rho"]), weights=True, solver=solver)
    assert namb == 0
    np.testing.assert_allclose(xa[..., sel], g["analysis"], rtol=RTOL, atol=ATOL)
    nw = g["weights"].shape[0]
    np.testing.assert_allclose(w[sel[:nw]], g["weights"], rtol=RTOL, atol=ATOL)
    _check_lists(eng, g["csr_off"], g["csr_idx"], sel=sel, w_ref=g["csr_w"])


@pytest.mark.parametrize("solver", ["newton", "jacobi"])
@pytest.mark.parametrize("k,n_grid,stride,radius", [(2, 40, 1, 3.0), (3, 64, 1, 2.5), (8, 300, 3, 4.0), (16, 200, 1, 9.0),
                                                    (23, 500, 2, 7.0), (32, 257, 1, 30.0), (40, 300, 2, 20.0),
                                                    (56, 200, 1, 10.0), (57, 200, 1, 10.0), (64, 300, 1, 12.0),
                                                    (72, 150, 1, 15.0), (80, 120, 1, 11.0), (88, 100, 1, 9.0),
                                                    (100, 96, 1, 6.0), (111, 64, 1, 8.0)])
def test_ensemble_sizes_against_oracle(k, n_grid, stride, radius, solver):
    """Every kernel configuration (tiles 1..14; 1/2/4/8 Gram warps per grid point; 1/2/4 solve warps per matrix) against
    the oracle, for both ensemble-space solvers."""
    m = _metrics()
    data = syn.lorenz96_1d(n_grid, k, stride, seed=100 + k)
    eng, xa, w, namb = _run(data, m.PeriodicDistance1D(float(n_grid)), radius, 1.07, weights=True, solver=solver)
    ref, wref, lists = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                          data["obs_rows"], orc.make_dist_periodic1d(float(n_grid)), radius,
                                          inf_factor=1.07, return_lists=True)
    np.testing.assert_allclose(xa, ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(w, wref, rtol=RTOL, atol=ATOL)
    off = np.zeros(n_grid + 1, dtype=np.int64); off[1:] = np.cumsum([len(l) for l in lists])
    _check_lists(eng, off, np.concatenate(lists).astype(np.int32))


def test_no_observation_in_reach_gives_inflated_prior():
    """core/etkf.py:91-95 per grid point: grid points without local observations get W = sqrt(rho) I."""
    m = _metrics()
    data = syn.lorenz96_1d(200, 12, 1, seed=7)
    keep = data["obs_rows"][:, 1] < 50
    data["obs_rows"] = data["obs_rows"][keep]
    data["normed_perts"] = np.ascontiguousarray(data["normed_perts"][:, keep]); data["normed_obs"] = data["normed_obs"][keep]
    eng, xa, w, _ = _run(data, m.AbsDistance1D(), 3.0, 1.21, weights=True)
    np.testing.assert_allclose(w[120], 1.1 * np.eye(12), atol=1e-13)
    ref, _ = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                data["obs_rows"], orc.dist_abs1d, 3.0, inf_factor=1.21)
    np.testing.assert_allclose(xa, ref, rtol=RTOL, atol=ATOL)
    # no observations at all
    eng.bin_obs(np.zeros((0, 1)), np.zeros((12, 0)), np.zeros((0,)))
    xa0 = eng.analyse(torch.as_tensor(data["state"].reshape(1, 12, 200)).cuda()).cpu().numpy()
    mean = data["state"].mean(axis=2, keepdims=True)
    np.testing.assert_allclose(xa0.reshape(data["state"].shape), mean + 1.1 * (data["state"] - mean), rtol=1e-12, atol=1e-12)


def test_host_entry_point_matches_device_path():
    """b200da_letkf_host (host buffers in, host buffers out) == device-resident path."""
    m = _metrics()
    data = syn.sphere_latlon(12, 24, 20, 800, seed=9, n_slices=2)
    st = data["state"]
    eng, xa, _, _ = _run(data, m.HaversineDistance(6371.0), 1500.0, 1.1)
    out = eng.analyse_host(st.reshape(2, 20, -1), data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
    np.testing.assert_array_equal(out.reshape(st.shape), xa)
    ref, _ = orc.letkf_analysis(st, data["normed_perts"], data["normed_obs"], data["grid_rows"], data["obs_rows"],
                                orc.make_dist_haversine(6371.0), 1500.0, inf_factor=1.1)
    np.testing.assert_allclose(xa, ref, rtol=RTOL, atol=ATOL)


def test_block_sharding_pack_unpack():
    """The multi-GPU decomposition on one device: analyse two block ranges separately, pack, unpack."""
    m = _metrics()
    data = syn.lorenz96_1d(333, 20, 2, seed=11)
    eng, xa, _, _ = _run(data, m.PeriodicDistance1D(333.0), 6.0, 1.05)
    x = torch.as_tensor(data["state"].reshape(1, 20, 333)).cuda()
    nb = eng.n_blocks
    half = nb // 2
    out_a = torch.zeros_like(x); out_b = torch.zeros_like(x)
    eng.analyse(x, out=out_a, blocks=(0, half)); eng.analyse(x, out=out_b, blocks=(half, nb))
    merged = torch.full_like(x, float("nan"))
    eng.unpack_columns(eng.pack_columns(out_a, 0, half), 0, half, merged)
    eng.unpack_columns(eng.pack_columns(out_b, half, nb), half, nb, merged)
    np.testing.assert_array_equal(merged.cpu().numpy().reshape(xa.shape), xa)


def test_full_size_properties_cfg2():
    """BASELINE cfg2 at full size (N=100k, k=40): size-independent properties instead of the oracle —
    (i) translation invariance on the ring, (ii) zero innovation keeps the ensemble mean,
    (iii) a sample of grid points against the oracle."""
    m = _metrics()
    data = syn.lorenz96_1d(100_000, 40, 2, seed=42)
    eng, xa, _, namb = _run(data, m.PeriodicDistance1D(100_000.0), 20.0, 1.1)
    assert namb == 0 and np.isfinite(xa).all()
    sel = np.arange(0, 100_000, 4999)
    ref, _ = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                data["obs_rows"], orc.make_dist_periodic1d(100_000.0), 20.0, inf_factor=1.1,
                                grid_subset=sel)
    np.testing.assert_allclose(xa[..., sel], ref, rtol=RTOL, atol=ATOL)
    # (ii) zero innovations: analysis mean == background mean
    data0 = dict(data); data0["normed_obs"] = np.zeros_like(data["normed_obs"])
    _, xa0, _, _ = _run(data0, m.PeriodicDistance1D(100_000.0), 20.0, 1.0)
    np.testing.assert_allclose(xa0.mean(axis=2), data["state"].mean(axis=2), rtol=1e-11, atol=1e-11)
    # (i) rolling grid + observations by 2 positions rolls the analysis
    rolled = dict(data)
    rolled["state"] = np.roll(data["state"], 2, axis=-1)
    rolled["normed_perts"] = np.ascontiguousarray(np.roll(data["normed_perts"], 1, axis=-1))
    rolled["normed_obs"] = np.roll(data["normed_obs"], 1)
    _, xar, _, _ = _run(rolled, m.PeriodicDistance1D(100_000.0), 20.0, 1.1)
    np.testing.assert_allclose(np.roll(xar, -2, axis=-1), xa, rtol=1e-11, atol=1e-11)


def test_analysis_is_deterministic():
    """Two runs on the same inputs are bit-identical (no order-dependent reductions, no races): 20k grid points on the
    sphere so that many CTAs of both kernels are in flight."""
    m = _metrics()
    data = syn.sphere_latlon(100, 200, 50, 20_000, seed=5)
    _, xa1, w1, _ = _run(data, m.HaversineDistance(6371.0), 1000.0, 1.1, weights=True)
    _, xa2, w2, _ = _run(data, m.HaversineDistance(6371.0), 1000.0, 1.1, weights=True)
    np.testing.assert_array_equal(xa1, xa2)
    np.testing.assert_array_equal(w1, w2)
    sel = np.arange(0, 20_000, 1999)
    ref, _ = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                data["obs_rows"], orc.make_dist_haversine(6371.0), 1000.0, inf_factor=1.1, grid_subset=sel)
    np.testing.assert_allclose(xa1[..., sel], ref, rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
def test_global_etkf_cfg4_shape_against_oracle(dtype, tol):
    """BASELINE cfg4 shape, scaled: global ETKF with k = 100 members, many observations, a long state (the split-M Gram,
    the single ensemble-space solve and the streaming update) against the oracle (interface/etkf.py:99-120, base.py:257-278)."""
    from pytassim_b200.engine import LETKFEngine
    m = _metrics()
    rng = np.random.RandomState(4)
    k, n, nobs = 100, 40_000, 3_000
    st = rng.normal(size=(1, 1, k, n))
    hx = st[0, 0][:, ::13][:, :nobs]
    yn = np.ascontiguousarray(hx - hx.mean(axis=0, keepdims=True))
    d = rng.normal(size=nobs) * 0.5
    npd = np.float64 if dtype == torch.float64 else np.float32
    ref, wref = orc.etkf_analysis(st.astype(npd).astype(np.float64), yn.astype(npd).astype(np.float64),
                                  d.astype(npd).astype(np.float64), inf_factor=1.1)
    eng = LETKFEngine(k, 1, m.AbsDistance1D(), 1.0, inf_factor=1.1, dtype=dtype)
    w = eng.etkf_weights(yn, d)
    xa = eng.apply_weights(torch.as_tensor(st.reshape(1, k, n), dtype=dtype).cuda(), w).cpu().numpy().reshape(st.shape)
    scale = np.abs(ref).max()
    assert np.abs(w.cpu().numpy() - wref).max() <= tol * max(1.0, np.abs(wref).max())
    assert np.abs(xa - ref).max() <= tol * scale


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
def test_sharded_etkf_entry_points(dtype, tol):
    """Observation- and state-sharded global ETKF on one GPU playing three ranks in turn: Grams of ragged
    observation ranges add up to the whole Gram, weights from the summed Gram equal b200da_etkf_weights, column-range
    updates tile the whole update bit-exactly, and everything agrees with the oracle."""
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200.parallel import ShardedETKF
    m = _metrics()
    rng = np.random.RandomState(14)
    k, n, nobs, n_sl = 50, 10_007, 2_501, 2
    npd = np.float64 if dtype == torch.float64 else np.float32
    st = rng.normal(size=(n_sl, 1, k, n)).astype(npd)
    hx = rng.normal(size=(k, nobs))
    yn = np.ascontiguousarray(hx - hx.mean(axis=0, keepdims=True)).astype(npd)
    d = (rng.normal(size=nobs) * 0.5).astype(npd)
    ref, wref = orc.etkf_analysis(st.astype(np.float64), yn.astype(np.float64), d.astype(np.float64), inf_factor=1.1)
    eng = LETKFEngine(k, n_sl, m.AbsDistance1D(), 1.0, inf_factor=1.1, dtype=dtype)
    ynd, dd = torch.as_tensor(yn).cuda(), torch.as_tensor(d).cuda()
    x = torch.as_tensor(st.reshape(n_sl, k, n)).cuda()
    w_whole = eng.etkf_weights(ynd, dd)
    xa_whole = eng.apply_weights(x, w_whole)
    sh = ShardedETKF(eng)
    sh.world = 3
    gram = torch.zeros((k + 1, k + 1), dtype=torch.float64, device="cuda")
    for j0, j1 in sh.ranges(nobs):
        gram += eng.etkf_gram(ynd, dd, obs_range=(j0, j1))
    aug = np.concatenate([yn.astype(np.float64), d.astype(np.float64)[None]], axis=0)
    gref = np.tril(aug @ aug.T)                       # element (k, k) = d d^T rides along for the kernelised ETKF
    assert np.abs(gram.cpu().numpy() - gref).max() <= 1e-12 * np.abs(gref).max()
    assert np.array_equal(np.triu(gram.cpu().numpy(), 1), np.zeros((k + 1, k + 1)))
    w = eng.etkf_weights_from_gram(gram, nobs)
    assert np.abs(w.cpu().numpy() - wref).max() <= tol * max(1.0, np.abs(wref).max())
    assert np.abs((w - w_whole).cpu().numpy()).max() <= tol * 0.1 * max(1.0, np.abs(wref).max())
    # one "rank": the whole range through the sharded entry points is bit-identical to the unsharded call
    g1 = eng.etkf_gram(ynd, dd)
    assert torch.equal(eng.etkf_weights_from_gram(g1, nobs), w_whole)
    out = torch.full_like(x, float("nan"))
    for c0, c1 in sh.ranges(n):
        eng.apply_weights_cols(x, w_whole, c0, c1, out)
    assert torch.equal(out, xa_whole)
    assert np.abs(out.cpu().numpy().reshape(st.shape) - ref).max() <= tol * np.abs(ref).max()
    # odd column boundaries (unaligned rows take the scalar stores) and per-grid weights
    out2 = torch.full_like(x, float("nan"))
    for c0, c1 in ((0, 333), (333, 4001), (4001, n)):
        eng.apply_weights_cols(x, w_whole, c0, c1, out2)
    assert torch.equal(out2, xa_whole)
    # the update fused with the gather: the kernel stores the same columns into further ("peer") analysis arrays
    out_a, out_b, out_c = (torch.full_like(x, float("nan")) for _ in range(3))
    for c0, c1 in ((0, 333), (333, 4001), (4001, n)):
        eng.apply_weights_cols(x, w_whole, c0, c1, out_a, peers=[out_b, out_c])
    assert torch.equal(out_a, xa_whole) and torch.equal(out_b, xa_whole) and torch.equal(out_c, xa_whole)
    wg = w_whole[None].repeat(n, 1, 1).contiguous()
    with pytest.raises(ValueError):
        eng.apply_weights_cols(x, wg, 0, n, out_a, peers=[out_b])
    out3 = torch.full_like(x, float("nan"))
    for c0, c1 in ((0, 5000), (5000, n)):
        eng.apply_weights_cols(x, wg, c0, c1, out3)
    assert torch.equal(out3, eng.apply_weights(x, wg))
    # no observations: the inflated prior (core/etkf.py:91-95)
    w0 = eng.etkf_weights_from_gram(torch.zeros_like(gram), 0).cpu().numpy()
    np.testing.assert_allclose(w0, np.sqrt(1.1) * np.eye(k), rtol=0, atol=1e-6 if dtype == torch.float32 else 1e-14)


@pytest.mark.parametrize("k,scale,tol", [(40, 1.0, 1e-10), (40, 10.0, 1e-10), (40, 30.0, 1e-10), (40, 100.0, 1e-10),
                                         (50, 300.0, 1e-10), (100, 200.0, 1e-10), (40, 1000.0, 2e-9), (24, 3000.0, 2e-8)])
def test_stiff_ensemble_space_problems_letkf(k, scale, tol):
    """Accurate / dense observations make C + (k-1)/rho I stiff (largest over smallest eigenvalue 1e3 ... 2e7 here).  The
    symmetric-storage Newton-Schulz iteration alone loses all accuracy beyond a ratio of a few thousand (measured: 1e-7 at
    1e4); the solve kernel must switch to its two-level form and stay within the FP64 tolerance of the north star.  For the
    last two cases (ratio 2e6 and 2e7) the eigendecomposition route of the reference is itself only reproducible to
    ~ratio * 1e-16 (the Jacobi solver and LAPACK differ by 6e-10 there), so the tolerance is widened accordingly."""
    m = _metrics()
    n_grid = 96
    data = syn.lorenz96_1d(n_grid, k, 1, seed=900 + k)
    data["normed_perts"] = data["normed_perts"] * scale                # observation error std 1 / scale
    data["normed_obs"] = data["normed_obs"] * scale
    eng, xa, w, namb = _run(data, m.PeriodicDistance1D(float(n_grid)), 6.0, 1.05, weights=True, solver="newton")
    ref, wref = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                   data["obs_rows"], orc.make_dist_periodic1d(float(n_grid)), 6.0, inf_factor=1.05)
    assert np.abs(w - wref).max() <= tol * max(1.0, np.abs(wref).max())
    assert np.abs(xa - ref).max() <= tol * np.abs(ref).max()


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
@pytest.mark.parametrize("k,nobs", [(100, 400_000), (50, 1_000_000), (128, 100_000)])
def test_stiff_global_etkf_weights(k, nobs, dtype, tol):
    """BASELINE cfg4 has 1e6 observations against a prior precision of (k-1)/rho: eigenvalue ratio ~1e4.  Weights against a
    numpy eigh solve of the device Gram (same arithmetic as core/etkf.py:57-77 on that Gram)."""
    from pytassim_b200.engine import LETKFEngine
    m = _metrics()
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    hx = torch.randn((k, nobs), dtype=torch.float64, device="cuda", generator=g)
    yn = (hx - hx.mean(dim=0, keepdim=True)).to(dtype).contiguous()
    d = (torch.randn(nobs, dtype=torch.float64, device="cuda", generator=g) * 0.5).to(dtype)
    rho = 1.1
    eng = LETKFEngine(k, 1, m.AbsDistance1D(), 1.0, inf_factor=rho, dtype=dtype)
    gram = eng.etkf_gram(yn, d).cpu().numpy()
    c = np.tril(gram[:k, :k]) + np.tril(gram[:k, :k], -1).T
    ev, u = np.linalg.eigh(c)
    ev = np.clip(ev, 0.0, None) + (k - 1) / rho
    assert ev.max() / ev.min() > 5e2
    wref = ((u / ev) @ u.T @ gram[k, :k])[:, None] + (u * np.sqrt((k - 1) / ev)) @ u.T
    w = eng.etkf_weights(yn, d).cpu().numpy().astype(np.float64)
    assert np.abs(w - wref).max() <= tol * np.abs(wref).max()


@pytest.mark.parametrize("k,n_grid,radius", [(112, 80, 7.0), (120, 72, 6.0), (128, 64, 9.0)])
def test_large_ensembles_newton_against_oracle(k, n_grid, radius):
    """Ensemble sizes beyond the shared-memory Jacobi limit (k > 111) run on the tensor-core Newton-Schulz solver only
    (BASELINE cfg5 goes up to k = 128)."""
    m = _metrics()
    data = syn.lorenz96_1d(n_grid, k, 1, seed=300 + k)
    eng, xa, w, namb = _run(data, m.PeriodicDistance1D(float(n_grid)), radius, 1.05, weights=True, solver="newton")
    ref, wref = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                   data["obs_rows"], orc.make_dist_periodic1d(float(n_grid)), radius, inf_factor=1.05)
    np.testing.assert_allclose(xa, ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(w, wref, rtol=RTOL, atol=ATOL)


def test_multi_row_localization_against_reference_golden(golden):
    """A dist_func with two rows (periodic ring distance x |level difference|) and two length scales
    (localization/gaspari_cohn.py:124-135): local-observation lists bit-exact, weights and analysis within 1e-10 of the
    reference's own GaspariCohn + hot loop (tests/golden/product_loc.npz)."""
    from pytassim_b200.engine import LETKFEngine
    m = _metrics()
    g = golden("product_loc.npz")
    k, n = g["state"].shape[2], g["state"].shape[3]
    metric = m.ProductDistance(m.PeriodicDistance1D(float(g["period"])), n_extra=1)
    eng = LETKFEngine(k, 1, metric, g["radius"], inf_factor=float(g["rho"]))
    eng.set_grid(g["grid_rows"][:, 1:])
    eng.bin_obs(g["obs_rows"][:, 1:], g["perts"], g["innov"])
    xa, w, namb = eng.analyse(torch.as_tensor(g["state"].reshape(1, k, n)).cuda(), return_weights=True, count_ambiguous=True)
    assert namb == 0
    np.testing.assert_allclose(w.cpu().numpy(), g["weights"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(xa.cpu().numpy().reshape(g["state"].shape), g["analysis"], rtol=RTOL, atol=ATOL)
    off, idx, wl, amb, _ = eng.neighbour_lists()
    np.testing.assert_array_equal(off.cpu().numpy(), g["csr_off"])
    np.testing.assert_array_equal(idx.cpu().numpy(), g["csr_idx"])
    wref = np.concatenate([g["w"][i][g["use"][i]] for i in range(n)])
    np.testing.assert_allclose(wl.cpu().numpy(), wref, rtol=0, atol=1e-14)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
def test_product_metric_sphere_with_levels_against_oracle(dtype, tol):
    """Horizontal haversine x vertical |dz| x time |dt| localization (three rows, three length scales) on the sphere, k = 50:
    index lists bit-exact, analysis within tolerance; FP32 plans run on the DMMA Gram (the tcgen05 Gram has one distance)."""
    from pytassim_b200.engine import LETKFEngine
    m = _metrics()
    data = syn.sphere_latlon(40, 80, 50, 6000, seed=31)
    rng = np.random.RandomState(32)
    n, mo = data["grid_rows"].shape[0], data["obs_rows"].shape[0]
    grid_rows = np.column_stack([data["grid_rows"], rng.randint(0, 5, n).astype(np.float64), np.zeros(n)])
    obs_rows = np.column_stack([data["obs_rows"], rng.uniform(0, 4, mo), rng.uniform(-3, 3, mo)])
    radius = (2500.0, 2.0, 4.0)
    npd = np.float64 if dtype == torch.float64 else np.float32
    st, yp, yo = data["state"].astype(npd), data["normed_perts"].astype(npd), data["normed_obs"].astype(npd)
    metric = m.ProductDistance(m.HaversineDistance(6371.0), n_extra=2)
    eng = LETKFEngine(50, 1, metric, radius, inf_factor=1.1, dtype=dtype)
    assert "tcgen05" not in eng.kernel_name
    eng.set_grid(grid_rows[:, 1:]); eng.bin_obs(obs_rows[:, 1:], yp, yo)
    xa = eng.analyse(torch.as_tensor(st.reshape(1, 50, n)).cuda()).cpu().numpy().reshape(st.shape)
    sel = np.arange(0, n, 97)
    dist = orc.make_dist_product(orc.make_dist_haversine(6371.0), 2, 2)
    ref, _, lists = orc.letkf_analysis(st.astype(np.float64), yp.astype(np.float64), yo.astype(np.float64), grid_rows, obs_rows,
                                       dist, radius, inf_factor=1.1, grid_subset=sel, return_lists=True)
    assert np.abs(xa[..., sel] - ref).max() <= tol * np.abs(ref).max()
    off, idx, _, _, namb = eng.neighbour_lists(with_weights=False)
    off, idx = off.cpu().numpy(), idx.cpu().numpy()
    for gi, l in zip(sel, lists):
        l = l[0] if isinstance(l, tuple) else l
        np.testing.assert_array_equal(idx[off[gi]:off[gi + 1]], l)
    assert len(lists[0][0] if isinstance(lists[0], tuple) else lists[0]) < mo // 4          # the extra rows do localize
    with pytest.raises(NotImplementedError):                                  # GaspariCohnInf evaluates a single distance
        LETKFEngine(50, 1, metric, radius, taper="gcinf")
    with pytest.raises(IndexError):                                           # fewer length scales than distance rows
        LETKFEngine(50, 1, metric, (2500.0,))


def test_error_paths_through_the_c_abi():
    """Status codes -> the reference's exception conventions: unsupported ensemble size, non-finite
    coordinates, call order, size mismatch (core/base.py:33-38), coordinate-count mismatch."""
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200 import _cabi
    m = _metrics()
    with pytest.raises(NotImplementedError):
        LETKFEngine(200, 1, m.AbsDistance1D(), 1.0)                     # k > 135: beyond the implemented tile counts
    with pytest.raises(ValueError):
        LETKFEngine(1, 1, m.AbsDistance1D(), 1.0)                       # an ensemble of one has no perturbations
    with pytest.raises(ValueError):
        LETKFEngine(8, 1, m.AbsDistance1D(), -1.0)
    eng = LETKFEngine(8, 1, m.EuclideanDistance(2), 1.0)
    with pytest.raises(ValueError):
        eng.set_grid(np.zeros((5, 3)))                                  # metric expects 2 coordinate columns
    bad = np.zeros((5, 2)); bad[2, 1] = np.nan
    with pytest.raises(ValueError):
        eng.set_grid(bad)                                               # non-finite coordinates
    with pytest.raises(_cabi.B200DAError):
        eng.bin_obs(np.zeros((3, 2)), np.zeros((8, 3)), np.zeros(3))    # bin_obs before a valid set_grid
    eng.set_grid(np.random.RandomState(0).uniform(0, 4, size=(5, 2)))
    with pytest.raises(ValueError, match="do not match"):
        eng.bin_obs(np.zeros((3, 2)), np.zeros((8, 3)), np.zeros(4))
    with pytest.raises(ValueError):
        eng.bin_obs(np.zeros((3, 2)), np.zeros((7, 3)), np.zeros(3))    # wrong ensemble size
    with pytest.raises(_cabi.B200DAError):
        eng.analyse(torch.zeros((1, 8, 5), dtype=torch.float64, device="cuda"))     # analyse before bin_obs
