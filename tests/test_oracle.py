"""The CPU oracle against the golden vectors produced by the reference's own leaf modules
(oracle/make_golden.py) and against the known-answer values in the reference's unit tests."""
import numpy as np
import pytest

import letkf_oracle as orc
from pytassim_b200.testing import synthetic as syn

RTOL = 1e-11


def test_core_kat_values(golden):
    """tests/unit_tests/core/test_etkf.py:142-210: cov = [[.75,.25],[.25,.75]], w_mean = [0.1,-0.1]."""
    g = golden("core_kat.npz")
    w_mean, w_perts, cov = orc.etkf_estimate_weights(g["normed_perts"], g["normed_obs"], 1.0)
    np.testing.assert_allclose(cov, [[0.75, 0.25], [0.25, 0.75]], atol=1e-14)
    np.testing.assert_allclose(w_mean, [[0.1], [-0.1]], atol=1e-14)
    np.testing.assert_allclose(w_perts @ w_perts.T, [[0.75, 0.25], [0.25, 0.75]], atol=1e-14)
    np.testing.assert_allclose(w_mean, g["w_mean"], atol=1e-14)
    np.testing.assert_allclose(w_perts, g["w_perts"], atol=1e-14)
    np.testing.assert_allclose(orc.etkf_weights(g["normed_perts"], g["normed_obs"], 1.0), g["W"], atol=1e-14)


def test_core_w_perts_is_eigh_formula(golden):
    """tests/unit_tests/core/test_etkf.py:168-179."""
    g = golden("core_random.npz")
    Y = g["Y0"]
    evals, evects = np.linalg.eigh(Y @ Y.T)
    evals = evals.clip(0) + (Y.shape[0] - 1)
    w_pert = evects @ np.diagflat(np.sqrt((Y.shape[0] - 1) / evals)) @ evects.T
    np.testing.assert_allclose(orc.etkf_estimate_weights(Y, g["d0"][None], 1.0)[1], w_pert, atol=1e-12)


@pytest.mark.parametrize("n", range(8))
def test_core_random_vs_reference(golden, n):
    g = golden("core_random.npz")
    W = orc.etkf_weights(g[f"Y{n}"], g[f"d{n}"], float(g[f"rho{n}"]))
    np.testing.assert_allclose(W, g[f"W{n}"], rtol=RTOL, atol=1e-12)
    # mean_j(W - I) = w_mean   (tests/unit_tests/core/test_etkf.py:219-225; holds for inf_factor = 1,
    # where the ones vector is an eigenvector of W_perts with eigenvalue 1)
    if float(g[f"rho{n}"]) == 1.0:
        w_mean = orc.etkf_estimate_weights(g[f"Y{n}"], g[f"d{n}"][None], 1.0)[0]
        np.testing.assert_allclose((W - np.eye(W.shape[0])).mean(axis=1), w_mean[:, 0], atol=1e-12)


def test_core_empty_obs_gives_inflated_prior(golden):
    """tests/unit_tests/core/test_etkf.py:227-233."""
    W = orc.etkf_weights(np.ones((10, 0)), np.ones((1, 0)), 1.1)
    np.testing.assert_allclose(W, np.sqrt(1.1) * np.eye(10), atol=1e-15)
    np.testing.assert_allclose(W, golden("core_random.npz")["W_empty"], atol=1e-15)


def test_core_size_mismatch_raises():
    """tests/unit_tests/core/test_etkf.py:235-240."""
    with pytest.raises(ValueError):
        orc.etkf_weights(np.ones((10, 4)), np.ones((1, 3)), 1.0)


def test_gaspari_cohn_pieces_bit_equal(golden):
    """tests/unit_tests/localization/test_gaspari_cohn.py:52-71,113-160 use assert_equal."""
    g = golden("gaspari_cohn.npz")
    with np.errstate(all="ignore"):
        np.testing.assert_equal(orc.gc_f1(g["dist"]), g["gc_f1"])
        np.testing.assert_equal(orc.gc_f2(g["dist"]), g["gc_f2"])
        np.testing.assert_equal(orc.gcinf_f1(g["dist"]), g["gci_f1"])
        np.testing.assert_equal(orc.gcinf_f2(g["dist"]), g["gci_f2"])
        np.testing.assert_equal(orc.gcinf_f3(g["dist"]), g["gci_f3"])
        np.testing.assert_equal(orc.gcinf_f4(g["dist"]), g["gci_f4"])


@pytest.mark.parametrize("grid_ind", [0, 10, 9999999])
def test_gaspari_cohn_localize_bit_equal(golden, grid_ind):
    g = golden("gaspari_cohn.npz")
    grid = np.arange(40.0)
    use, w = orc.gaspari_cohn_localize(np.abs(grid_ind - grid), 5.0)
    np.testing.assert_equal(use, g[f"gc_use_{grid_ind}"])
    np.testing.assert_equal(w, g[f"gc_w_{grid_ind}"])
    use, w = orc.gaspari_cohn_inf_localize(np.abs(grid_ind - grid), 5.0)
    np.testing.assert_equal(use, g[f"gci_use_{grid_ind}"])
    np.testing.assert_equal(w, g[f"gci_w_{grid_ind}"])
    if grid_ind == 9999999:      # test_gaspari_cohn.py:73-76: zero weight beyond 2 c
        assert not use.any() and not w.any()


def test_gaspari_cohn_dense_r_and_two_components(golden):
    g = golden("gaspari_cohn.npz")
    use, w = orc.gaspari_cohn_localize(g["r"], 1.0)
    np.testing.assert_equal(use, g["gc_use_r"]); np.testing.assert_equal(w, g["gc_w_r"])
    use, w = orc.gaspari_cohn_inf_localize(g["r"], 1.0)
    np.testing.assert_equal(use, g["gci_use_r"]); np.testing.assert_equal(w, g["gci_w_r"])
    grid = np.arange(40.0)
    use, w = orc.gaspari_cohn_localize((np.abs(10 - grid), np.abs(10 - grid) * 0.25), (5.0, 2.0))
    np.testing.assert_equal(use, g["gc2_use_10"]); np.testing.assert_equal(w, g["gc2_w_10"])


def _csr(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in lists])
    return off, np.concatenate(lists).astype(np.int32)


def test_fixture_letkf_gc10(golden):
    """tests/unit_tests/interface/test_letkf.py:106-157 on the reference fixtures (time slice 0)."""
    g = golden("fixture_letkf.npz")
    state = g["state"][:, :1]
    hx = np.transpose(state[0], (1, 0, 2))
    innov, perts = orc.obs_space_variables([hx], [g["obs"][:1]], [g["cov"]])
    np.testing.assert_allclose(innov, g["a_innov"], rtol=0, atol=0)
    ana, W, lists = orc.letkf_analysis(state, perts, innov, g["a_grid_rows"], g["a_obs_rows"], orc.dist_abs1d,
                                       (10.0,), return_lists=True)
    off, idx = _csr(lists)
    np.testing.assert_array_equal(off, g["a_csr_off"]); np.testing.assert_array_equal(idx, g["a_csr_idx"])
    np.testing.assert_allclose(W, g["a_weights"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(ana, g["a_analysis"], rtol=1e-10, atol=1e-10)
    # SURVEY.md 8(c) check values
    np.testing.assert_allclose(ana.sum(), -2.616127006116, atol=1e-9)
    np.testing.assert_allclose(ana[0, 0, 0, :3], [0.317924571007, 0.026543174387, 0.749034401420], atol=1e-10)
    assert [len(l) for l in lists[:12]] == list(range(20, 32))
    np.testing.assert_array_equal(lists[0], np.arange(20))


def test_fixture_global_etkf_two_datasets(golden):
    """tests/unit_tests/interface/test_letkf.py:64-70: LETKF without localization == ETKF, two obs datasets."""
    g = golden("fixture_letkf.npz")
    state = g["state"][:, 2:]
    hx = np.transpose(state[0], (1, 0, 2))
    innov, perts = orc.obs_space_variables([hx] * 2, [g["obs"][2:]] * 2, [g["cov"]] * 2)
    np.testing.assert_array_equal(innov, g["b_innov"]); np.testing.assert_array_equal(perts, g["b_perts"])
    ana, W = orc.etkf_analysis(state, perts, innov, 1.0)
    np.testing.assert_allclose(W, g["b_weights"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(ana, g["b_analysis"], rtol=1e-10, atol=1e-10)
    grid_rows = np.stack([np.zeros(40), np.arange(40.0)], axis=1)
    obs_rows = np.concatenate([grid_rows, grid_rows])
    ana_l, _ = orc.letkf_analysis(state, perts, innov, grid_rows, obs_rows, None, None)
    np.testing.assert_allclose(ana_l, ana, rtol=1e-10, atol=1e-10)
    # all-ones localization (test_letkf.py:94-104): dist == 0 everywhere
    ana_1, _ = orc.letkf_analysis(state, perts, innov, grid_rows, obs_rows,
                                  lambda x, y: np.zeros(y.shape[0]), (1.0, 1.0))
    np.testing.assert_allclose(ana_1, ana, rtol=1e-10, atol=1e-10)


def test_fixture_letkf_gcinf(golden):
    g = golden("fixture_letkf.npz")
    state = g["state"][:, 1:2]
    ana, W, lists = orc.letkf_analysis(state, g["c_perts"], g["c_innov"], g["a_grid_rows"], g["a_obs_rows"],
                                       orc.dist_abs1d, 8.0, inf_factor=1.1, taper="gcinf", return_lists=True)
    off, idx = _csr(lists)
    np.testing.assert_array_equal(off, g["c_csr_off"]); np.testing.assert_array_equal(idx, g["c_csr_idx"])
    np.testing.assert_allclose(ana, g["c_analysis"], rtol=1e-10, atol=1e-10)


def test_mul_rcinv_chol_equals_diag():
    """observation.py:247-275: correlated form with a diagonal R equals value / sqrt(var)."""
    rnd = np.random.RandomState(0)
    v = rnd.normal(size=(3, 5)); var = rnd.uniform(0.5, 2, size=5)
    np.testing.assert_allclose(orc.mul_rcinv(v, np.diag(var)), orc.mul_rcinv(v, var), rtol=1e-14)


def test_apply_weights_identity_and_formula():
    """tests/unit_tests/interface/test_base.py:330-349."""
    rnd = np.random.RandomState(1)
    st = rnd.normal(size=(2, 3, 5, 7)); W = rnd.normal(size=(7, 5, 5))
    np.testing.assert_allclose(orc.apply_weights(st, np.eye(5)), st, atol=1e-14)
    mean = st.mean(axis=2, keepdims=True); p = st - mean
    right = mean + np.stack([np.stack([(p[..., :, g][..., :, None] * W[g]).sum(axis=-2) for g in range(7)], -1)], 0)[0]
    np.testing.assert_allclose(orc.apply_weights(st, W), right, atol=1e-13)


SYNTH = {
    "cfg1_l96_n40_k50.npz": lambda g: (syn.lorenz96_1d(40, 50, 1, seed=42), orc.make_dist_periodic1d(40.0)),
    "cfg2_l96_n2000_k40.npz": lambda g: (syn.lorenz96_1d(2000, 40, 2, seed=43), orc.make_dist_periodic1d(2000.0)),
    "bench_default_n1000_k50.npz": lambda g: (syn.lorenz96_1d(1000, 50, 10, seed=45), orc.dist_abs1d),
    "cfg3_sphere_small.npz": lambda g: (syn.sphere_latlon(24, 48, 50, 3000, seed=44), orc.make_dist_haversine(6371.0)),
    "euclid2d_k16.npz": lambda g: ({k: g[k] for k in ("state", "normed_perts", "normed_obs", "grid_rows", "obs_rows")},
                                   orc.dist_euclid),
}


@pytest.mark.parametrize("name", sorted(SYNTH))
def test_synthetic_configs_vs_reference(golden, name):
    g = golden(name)
    data, dist = SYNTH[name](g)
    sel = g["sel"]
    ana, W, lists = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                       data["obs_rows"], dist, float(g["radius"]), inf_factor=float(g["rho"]),
                                       grid_subset=sel, return_lists=True)
    off, idx = _csr(lists)
    np.testing.assert_array_equal(off, g["csr_off"])
    np.testing.assert_array_equal(idx, g["csr_idx"])
    np.testing.assert_allclose(W[:g["weights"].shape[0]], g["weights"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(ana, g["analysis"], rtol=1e-10, atol=1e-10)


def test_ketkf_linear_restatement_against_reference(golden):
    """core/ketkf.py:69-100 with LinearKernel: the restatement reproduces the reference module on centred and uncentred
    perturbations; on centred perturbations (what the interface always hands over, base.py:367-372) it equals the ETKF core,
    which is why KETKF / LKETKF run on the ETKF device path."""
    g = golden("ketkf_linear.npz")
    for i in range(int(g["n_cases"])):
        perts, raw, obs, rho = g["c%d_perts" % i], g["c%d_raw" % i], g["c%d_obs" % i], float(g["c%d_rho" % i])
        np.testing.assert_allclose(orc.ketkf_linear_weights(perts, obs, rho), g["c%d_w" % i], rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(orc.ketkf_linear_weights(raw, obs, rho), g["c%d_w_raw" % i], rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(orc.etkf_weights(perts, obs, rho), g["c%d_w" % i], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(orc.ketkf_linear_weights(np.zeros((6, 0)), np.zeros((1, 0)), 1.3), g["empty_w"], rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        orc.ketkf_linear_weights(np.zeros((3, 4)), np.zeros((1, 5)))
    # localized KETKF on the reference fixtures == localized ETKF restatement
    grid_rows = np.stack([np.zeros(40), g["lketkf_grid"]], axis=1)
    obs_rows = np.stack([np.zeros(40), g["lketkf_obs_grid"]], axis=1)
    ana, w = orc.letkf_analysis(g["lketkf_state"], g["lketkf_perts"], g["lketkf_innov"], grid_rows, obs_rows,
                                orc.dist_abs1d, 10.0, inf_factor=1.1)
    np.testing.assert_allclose(w, g["lketkf_weights"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(ana, g["lketkf_analysis"], rtol=1e-11, atol=1e-12)


def test_multi_row_localization_against_reference(golden):
    """gaspari_cohn.py:124-135 with a dist_func that returns two rows and two length scales: mask, weights, index lists and the
    LETKF analysis of the restatement against the reference's own GaspariCohn + hot loop (tests/golden/product_loc.npz)."""
    g = golden("product_loc.npz")
    dist = orc.make_dist_product(orc.make_dist_periodic1d(float(g["period"])), 1, 1)
    for gi in range(0, g["grid_rows"].shape[0], 5):
        use, w = orc.gaspari_cohn_localize(dist(g["grid_rows"][gi], g["obs_rows"]), g["radius"])
        np.testing.assert_array_equal(use, g["use"][gi])
        np.testing.assert_array_equal(w, g["w"][gi])
    ana, w, lists = orc.letkf_analysis(g["state"], g["perts"], g["innov"], g["grid_rows"], g["obs_rows"], dist, g["radius"],
                                       inf_factor=float(g["rho"]), return_lists=True)
    np.testing.assert_allclose(ana, g["analysis"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(w, g["weights"], rtol=1e-12, atol=1e-12)
    idx = np.concatenate([l[0] if isinstance(l, tuple) else l for l in lists])
    np.testing.assert_array_equal(idx, g["csr_idx"])


def test_ketkf_kernel_restatements_against_reference(golden):
    """core/ketkf.py:69-100 with every kernel configuration of pytassim_b200/testing/kernel_cases.py: the numpy restatement
    of kernels/*.py reproduces ``KETKFModule(kernel)`` of the reference (tests/golden/ketkf_kernels.npz, written by
    ``oracle/make_golden.py kernels`` from the reference's own modules), global and localized."""
    from pytassim_b200.testing import kernel_cases as kc
    g = golden("ketkf_kernels.npz")
    assert int(g["n_cases"]) == len(kc.PROBLEM_SIZES)
    for i, (k, p, rho) in enumerate(kc.PROBLEM_SIZES):
        perts, obs = g["c%d_perts" % i], g["c%d_obs" % i]
        assert perts.shape == (k, p) and float(g["c%d_rho" % i]) == rho
        for name, build in kc.KERNEL_CASES:
            w = orc.ketkf_weights(perts, obs, rho, build(orc, p))
            np.testing.assert_allclose(w, g["c%d_w_%s" % (i, name)], rtol=1e-11, atol=1e-12, err_msg=name)
    w = orc.ketkf_weights(g["c0_perts"], g["c0_obs"], 1.1, orc.OrnsteinUhlenbeckKernel(30.))
    np.testing.assert_allclose(w, g["c0_w_ornuhl"], rtol=1e-11, atol=1e-12)
    grid_rows = np.stack([np.zeros(40), g["lketkf_grid"]], axis=1)
    obs_rows = np.stack([np.zeros(40), g["lketkf_obs_grid"]], axis=1)
    for name, build in kc.KERNEL_CASES:
        ws = np.stack([orc.lketkf_weights_point(grid_rows[j], g["lketkf_perts"], g["lketkf_innov"][None], obs_rows,
                                                orc.dist_abs1d, (10.,), build(orc, 20), inf_factor=1.1) for j in range(40)])
        np.testing.assert_allclose(ws, g["lketkf_weights_" + name], rtol=1e-11, atol=1e-12, err_msg=name)
        np.testing.assert_allclose(orc.apply_weights(g["lketkf_state"], ws), g["lketkf_analysis_" + name], rtol=1e-11, atol=1e-12)


def test_ienks_restatement_against_reference(golden):
    """core/ienks.py:28-174: the numpy restatement reproduces ``IEnKSTransformModule`` / ``IEnKSBundleModule`` of the reference
    over three iterations (tests/golden/ienks.npz from ``oracle/make_golden.py ienks``), the hand-through without observations
    and the localized call with skipped weights (interface/lienks.py:68-118)."""
    g = golden("ienks.npz")
    for i in range(int(g["n_cases"])):
        perts, obs, tau = g["c%d_perts" % i], g["c%d_obs" % i], float(g["c%d_tau" % i])
        for variant, eps in (("transform", None), ("bundle", float(g["c%d_eps" % i]))):
            w = np.eye(perts.shape[0])
            for it in range(3):
                w = orc.ienks_weights(w, perts * (1.0 if eps is None else eps), obs, tau, eps)
                np.testing.assert_allclose(w, g["c%d_%s_w%d" % (i, variant, it)], rtol=1e-11, atol=1e-12)
    assert np.array_equal(orc.ienks_weights(g["empty_in"], np.zeros((6, 0)), np.zeros((1, 0)), 0.7), g["empty_w"])
    with pytest.raises(ValueError):
        orc.ienks_weights(np.eye(3), np.zeros((3, 4)), np.zeros((1, 5)))
    grid_rows = np.stack([np.zeros(40), g["l_grid"]], axis=1)
    obs_rows = np.stack([np.zeros(40), g["l_obs_grid"]], axis=1)
    for variant, eps in (("transform", None), ("bundle", float(g["l_eps"]))):
        tau = float(g["l_%s_tau" % variant])
        weights = np.stack([np.eye(10)] * 40)
        for it in range(3):
            weights = np.stack([orc.lienks_weights_point(grid_rows[j], weights[j], g["l_perts"] * (1.0 if eps is None else eps),
                                                         g["l_innov"][None], obs_rows, orc.dist_abs1d, (10.,), tau, eps)
                                for j in range(40)])
            np.testing.assert_allclose(weights, g["l_%s_w%d" % (variant, it)], rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(orc.apply_weights(g["l_state"], weights), g["l_%s_analysis" % variant], rtol=1e-11, atol=1e-12)
