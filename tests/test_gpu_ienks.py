"""IEnKS weight update (transform and bundle variants) on the GPU through the C ABI (b200da_etkf_ienks_weights,
b200da_letkf_ienks): Gram -> k_ienks_pre -> ensemble-space solve -> k_ienks_keep, against the reference's
``IEnKSTransformModule`` / ``IEnKSBundleModule`` (tests/golden/ienks.npz, generated from pytassim/core/ienks.py by
``oracle/make_golden.py ienks``) and against the oracle restatement.  Tolerance: FP64 rtol = atol = 1e-10."""
import numpy as np
import pytest
import torch

import letkf_oracle as orc
from pytassim_b200.engine import LETKFEngine
from pytassim_b200.localization import metrics as m
from pytassim_b200.testing import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-10, atol=1e-10)
VARIANTS = [("transform", None), ("bundle", 1e-2)]
TOL9 = 1e-10     # was 1e-9 in round 1


def _close(got, ref, tol):
    """assert_allclose(rtol = atol = tol) that also reports how much of the tolerance was used."""
    got, ref = np.asarray(got), np.asarray(ref)
    used = float(np.max(np.abs(got - ref) / (tol + tol * np.abs(ref))))
    print("tolerance used: {0:.3f} of {1:g}".format(used, tol))
    np.testing.assert_allclose(got, ref, rtol=tol, atol=tol)


@pytest.mark.parametrize("variant,eps", VARIANTS)
@pytest.mark.parametrize("solver", ["newton", "jacobi"])
def test_global_ienks_iterations_against_reference(golden, variant, eps, solver):
    """interface/ienks.py:96-118 -> core/ienks.py:134-174: three iterations from the prior identity, every iterate compared
    (the device gets the reference's previous iterate, so errors do not accumulate across the comparison)."""
    g = golden("ienks.npz")
    for i in range(int(g["n_cases"])):
        perts, obs, tau = g["c%d_perts" % i], g["c%d_obs" % i], float(g["c%d_tau" % i])
        k = perts.shape[0]
        scale = 1.0 if eps is None else eps
        eng = LETKFEngine(k, 1, m.AbsDistance1D(), 1.0).set_solver(solver)
        w = np.eye(k)
        for it in range(3):
            w_dev = eng.ienks_weights(w, perts * scale, obs.reshape(-1), tau=tau, epsilon=eps).cpu().numpy()
            ref = g["c%d_%s_w%d" % (i, variant, it)]
            np.testing.assert_allclose(w_dev, ref, err_msg="case %d iteration %d" % (i, it), **TOL)
            w = ref
        # chained on the device: three iterations without the reference in between
        w = torch.eye(k, dtype=torch.float64, device="cuda")
        for it in range(3):
            w = eng.ienks_weights(w, perts * scale, obs.reshape(-1), tau=tau, epsilon=eps)
        _close(w.cpu().numpy(), g["c%d_%s_w2" % (i, variant)], TOL9)
    # no observations: the weights are handed through (core/ienks.py:143)
    eng = LETKFEngine(6, 1, m.AbsDistance1D(), 1.0)
    w0 = eng.ienks_weights(g["empty_in"], np.zeros((6, 0)), np.zeros(0), tau=0.7).cpu().numpy()
    assert np.array_equal(w0, g["empty_w"])


@pytest.mark.parametrize("variant,eps", VARIANTS)
def test_localized_ienks_fixture_against_reference(golden, variant, eps):
    """interface/lienks.py:68-118 on the reference fixtures (GaspariCohn((10.,), |grid - obs|)): three iterations of the
    per-grid-point weights (first from the prior identity, then from the (N, k, k) weights) and the final analysis."""
    g = golden("ienks.npz")
    state, tau = g["l_state"], float(g["l_%s_tau" % variant])
    scale = 1.0 if eps is None else eps
    eng = LETKFEngine(10, 2, m.AbsDistance1D(), 10.0)
    eng.set_grid(g["l_grid"][:, None])
    eng.bin_obs(g["l_obs_grid"][:, None], g["l_perts"] * scale, g["l_innov"])
    x = torch.as_tensor(state.reshape(2, 10, 40)).cuda()
    w = np.eye(10)
    for it in range(3):
        xa, w_dev = eng.ienks_step(x, w, tau=tau, epsilon=eps)
        ref = g["l_%s_w%d" % (variant, it)]
        np.testing.assert_allclose(w_dev.cpu().numpy(), ref, err_msg="iteration %d" % it, **TOL)
        w = ref
    np.testing.assert_allclose(xa.cpu().numpy().reshape(state.shape), g["l_%s_analysis" % variant], **TOL)


@pytest.mark.parametrize("k,tau,eps", [(40, 1.0, None), (50, 0.7, None), (24, 0.5, 1e-2), (50, 1.0, 1e-3)])
def test_localized_ienks_ring_against_oracle(k, tau, eps):
    """Lorenz-96 ring (2000 grid points, every 2nd observed, a stretch without observations): two chained device iterations
    against the oracle on a subset of grid points; grid points without local observations keep their incoming weights and
    their state columns are updated with those (core/ienks.py:143, interface/base.py:257-278)."""
    n = 2000
    data = syn.lorenz96_1d(n, k, 2, seed=11)
    keep = ~((data["obs_rows"][:, 1] > 900) & (data["obs_rows"][:, 1] < 1100))
    obs_rows, innov = data["obs_rows"][keep], data["normed_obs"][keep]
    perts = data["normed_perts"][:, keep] * (1.0 if eps is None else eps)
    eng = LETKFEngine(k, 1, m.PeriodicDistance1D(float(n)), 20.0)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(obs_rows[:, 1:], perts, innov)
    x = torch.as_tensor(data["state"].reshape(1, k, n)).cuda()
    rng = np.random.RandomState(3)
    w_start = np.eye(k) + 0.05 * rng.normal(size=(k, k))          # a non-trivial incoming matrix, the same for every grid point
    _, w1 = eng.ienks_step(x, w_start, tau=tau, epsilon=eps)
    xa, w2 = eng.ienks_step(x, w1, tau=tau, epsilon=eps)
    sel = np.concatenate([np.arange(0, n, 41), np.arange(985, 1015)])
    dist = orc.make_dist_periodic1d(float(n))
    ref = []
    for j in sel:
        r1 = orc.lienks_weights_point(data["grid_rows"][j], w_start, perts, innov[None], obs_rows, dist, (20.,), tau, eps)
        ref.append(orc.lienks_weights_point(data["grid_rows"][j], r1, perts, innov[None], obs_rows, dist, (20.,), tau, eps))
    ref = np.stack(ref)
    _close(w2.cpu().numpy()[sel], ref, TOL9)
    assert np.array_equal(w2.cpu().numpy()[1000], w_start)        # no local observation: handed through twice
    _close(xa.cpu().numpy().reshape(1, 1, k, n)[..., sel], orc.apply_weights(data["state"][..., sel], ref),
                               TOL9)


def test_zero_gram_with_local_observations_is_handed_through():
    """The one accepted semantic difference: the IEnKS pre-pass recognises "no local observation"
    (core/ienks.py:143, the localized wrapper skips the module) by an all-zero Gram slot.  A grid point whose local observations
    EXIST but whose localized perturbations and innovations are all exactly zero therefore keeps its incoming weights, whereas
    the reference runs its update on the zero Gram.  Pinned here: (a) the device hands the weights through bit for bit,
    (b) with the prior identity weights (first iteration of every IEnKS run) that IS the reference's result, the identity being
    the fixed point of the update on a zero Gram, (c) with other incoming weights the reference's update differs - the
    difference this test documents."""
    n, k, tau = 64, 12, 0.7
    data = syn.lorenz96_1d(n, k, 1, seed=9)
    perts = np.zeros_like(data["normed_perts"])
    innov = np.zeros_like(data["normed_obs"])
    eng = LETKFEngine(k, 1, m.PeriodicDistance1D(float(n)), 4.0)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(data["obs_rows"][:, 1:], perts, innov)
    counts, _ = eng.neighbour_counts()
    assert int(counts.min()) > 0                                   # every grid point does have local observations
    x = torch.as_tensor(data["state"].reshape(1, k, n)).cuda()
    dist = orc.make_dist_periodic1d(float(n))
    # (a) + (b): identity in, identity out on both sides
    _, w_id = eng.ienks_step(x, np.eye(k), tau=tau)
    assert np.array_equal(w_id.cpu().numpy(), np.broadcast_to(np.eye(k), (n, k, k)))
    ref_id = orc.lienks_weights_point(data["grid_rows"][5], np.eye(k), perts, innov[None], data["obs_rows"], dist, (4.,), tau, None)
    np.testing.assert_allclose(ref_id, np.eye(k), rtol=0, atol=1e-13)
    # (a) + (c): a non-trivial incoming matrix is handed through by the device, moved by the reference
    w_in = np.eye(k) + 0.05 * np.random.RandomState(4).normal(size=(k, k))
    xa, w_out = eng.ienks_step(x, w_in, tau=tau)
    assert np.array_equal(w_out.cpu().numpy(), np.broadcast_to(w_in, (n, k, k)))
    np.testing.assert_allclose(xa.cpu().numpy().reshape(1, 1, k, n), orc.apply_weights(data["state"], np.broadcast_to(w_in, (n, k, k))),
                               rtol=1e-12, atol=1e-12)
    ref = orc.lienks_weights_point(data["grid_rows"][5], w_in, perts, innov[None], data["obs_rows"], dist, (4.,), tau, None)
    assert np.abs(ref - w_in).max() > 1e-3                         # the reference's update on the zero Gram is not a no-op


def test_ienks_argument_checks():
    eng = LETKFEngine(10, 1, m.AbsDistance1D(), 1.0)
    with pytest.raises(ValueError):
        eng.ienks_weights(np.eye(10), np.zeros((10, 4)), np.zeros(4), tau=0.0)        # tau in (0, 1]
    with pytest.raises(ValueError):
        eng.ienks_weights(np.eye(10), np.zeros((10, 4)), np.zeros(5))                 # core/base.py:28-38
    with pytest.raises(NotImplementedError):
        LETKFEngine(100, 1, m.AbsDistance1D(), 1.0).ienks_weights(np.eye(100), np.zeros((100, 4)), np.zeros(4))   # k <= 96


def _forward_model(st, iter_num):
    """A weakly non-linear 'model': (analysis state, pseudo state used for the observation equivalents)."""
    v = np.asarray(st.values)
    return st, st.copy(data=v + 0.05 * np.tanh(v))


@pytest.mark.parametrize("localized", [True, False])
@pytest.mark.parametrize("eps", [None, 1e-3])
def test_ienks_interface_classes_against_oracle_loop(golden, localized, eps):
    """interface/variational.py:105-135 through ``assimilate``: three outer iterations (model propagation of the weighted
    ensemble, observation operator, correlated-R normalisation, weight update, final update) on the reference fixtures
    against the same loop written with the oracle's functions."""
    from pytassim_b200.interface import IEnKSTransform, IEnKSBundle, LocalizedIEnKSTransform, LocalizedIEnKSBundle
    from pytassim_b200.localization import GaspariCohn, AbsDistance1D
    from test_host_logic import _fixture_objects
    g, state, obs = _fixture_objects(golden)
    ob0 = obs.isel(time=[0])
    tau, n_iter, k = 0.8, 3, 10
    loc = GaspariCohn((10.,), AbsDistance1D()) if localized else None
    if eps is None:
        alg = LocalizedIEnKSTransform(_forward_model, localization=loc, tau=tau, max_iter=n_iter) if localized \
            else IEnKSTransform(_forward_model, tau=tau, max_iter=n_iter)
    else:
        alg = LocalizedIEnKSBundle(_forward_model, localization=loc, tau=tau, epsilon=eps, max_iter=n_iter) if localized \
            else IEnKSBundle(_forward_model, tau=tau, epsilon=eps, max_iter=n_iter)
    ana = alg.assimilate(state, ob0, analysis_time="1992-12-25 00:00")
    assert ana.dims == state.dims and ana.shape == (2, 1, 10, 40)
    # the same loop with the oracle
    st0 = g["state"][:, :1]
    grid_rows = np.stack([np.zeros(40), g["grid"]], axis=1)
    obs_rows = np.stack([np.zeros(40), g["obs_grid"]], axis=1)
    weights = np.stack([np.eye(k)] * 40)
    for _ in range(n_iter):
        model_w = weights if eps is None else eps * np.eye(k) + weights.mean(axis=-1, keepdims=True)   # ienks.py:153-160
        model = orc.apply_weights(st0, model_w)
        pseudo = model + 0.05 * np.tanh(model)
        innov, perts = orc.obs_space_variables([pseudo[0, 0][:, None, :]], [g["obs"][:1]], [g["cov"]])
        if localized:
            weights = np.stack([orc.lienks_weights_point(grid_rows[j], weights[j], perts, innov[None], obs_rows, orc.dist_abs1d,
                                                         (10.,), tau, eps) for j in range(40)])
        else:
            weights = np.stack([orc.ienks_weights(weights[0], perts, innov[None], tau, eps)] * 40)
    ref = orc.apply_weights(st0, weights)
    _close(ana.values, ref, TOL9)


def test_ienks_smoother_and_weight_store(golden, tmp_path):
    """variational.py:132-133 (smoother mode propagates the analysis once more) and :56-80 (weights of every iteration go
    through the netCDF store when ``weight_save_path`` is set)."""
    from pytassim_b200.interface import LocalizedIEnKSTransform
    from pytassim_b200.localization import GaspariCohn, AbsDistance1D
    from test_host_logic import _fixture_objects
    g, state, obs = _fixture_objects(golden)
    ob0 = obs.isel(time=[0])
    calls = []

    def model(st, iter_num):
        calls.append(iter_num)
        return st.copy(data=np.asarray(st.values) * 2.0), st

    path = str(tmp_path / "ienks_weights.nc")
    kw = dict(localization=GaspariCohn((10.,), AbsDistance1D()), tau=1.0, max_iter=2)
    plain = LocalizedIEnKSTransform(model, **kw).assimilate(state, ob0, analysis_time="1992-12-25 00:00")
    assert calls == [0, 1]
    stored = LocalizedIEnKSTransform(model, weight_save_path=path, smoother=True, **kw)
    ana = stored.assimilate(state, ob0, analysis_time="1992-12-25 00:00")
    assert calls == [0, 1, 0, 1, 2]
    np.testing.assert_allclose(ana.values, 2.0 * plain.values, rtol=1e-12, atol=1e-12)
    w = stored.load_weights()
    assert w.dims == ('grid', 'ensemble', 'ensemble_new') and w.shape == (40, 10, 10)
    # with tau = 1 and an identity model the first iteration is the LETKF without inflation (core/ienks.py vs core/etkf.py)
    one = LocalizedIEnKSTransform(lambda st, it: (st, st), localization=kw["localization"], tau=1.0, max_iter=1)
    np.testing.assert_allclose(one.assimilate(state, ob0, analysis_time="1992-12-25 00:00").values, g["a_analysis"], **TOL)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-10), (torch.float32, 1e-4)])
def test_ienks_first_iteration_is_letkf_at_scale(dtype, tol):
    """Size-independent identity at the cfg2 size (N = 100 000, k = 40, every 2nd variable observed): from the prior identity
    weights with tau = 1 the transform IEnKS step is the LETKF without inflation (core/ienks.py:117-132 with W = I reduces to
    core/etkf.py:57-77), and the bundle step with perturbations scaled by epsilon is the same matrix.  FP32 plans take the
    tcgen05 Gram in front of the IEnKS pre-pass."""
    n, k = 100_000, 40
    data = syn.lorenz96_1d(n, k, 2, seed=42)
    npd = np.float64 if dtype == torch.float64 else np.float32
    eng = LETKFEngine(k, 1, m.PeriodicDistance1D(float(n)), 20.0, inf_factor=1.0, dtype=dtype)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"].astype(npd), data["normed_obs"].astype(npd))
    x = torch.as_tensor(data["state"].reshape(1, k, n).astype(npd)).cuda()
    xa_ref, w_ref = eng.analyse(x, return_weights=True)
    xa, w = eng.ienks_step(x, np.eye(k), tau=1.0)
    scale = float(w_ref.abs().max())
    assert float((w - w_ref).abs().max()) <= tol * scale
    assert float((xa - xa_ref).abs().max()) <= tol * float(xa_ref.abs().max())
    if dtype == torch.float64:
        eps = 2.0 ** -7                                             # a power of two: the scaling is exact
        eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"] * eps, data["normed_obs"])
        _, wb = eng.ienks_step(x, np.eye(k), tau=1.0, epsilon=eps)
        assert float((wb - w_ref).abs().max()) <= tol * scale


@pytest.mark.parametrize("eps", [None, 1e-3])
def test_localized_ienks_six_chained_iterations(eps):
    """Error growth over a chain: six localized iterations on the device without the oracle in between (every iteration
    inverts the previous weights, core/ienks.py:62-65) stay within 1e-9 of the oracle chain."""
    n, k, tau, n_iter = 600, 24, 0.9, 6
    data = syn.lorenz96_1d(n, k, 2, seed=19)
    perts = data["normed_perts"] * (1.0 if eps is None else eps)
    eng = LETKFEngine(k, 1, m.PeriodicDistance1D(float(n)), 20.0)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(data["obs_rows"][:, 1:], perts, data["normed_obs"])
    x = torch.as_tensor(data["state"].reshape(1, k, n)).cuda()
    w = torch.eye(k, dtype=torch.float64, device="cuda")
    for _ in range(n_iter):
        xa, w = eng.ienks_step(x, w, tau=tau, epsilon=eps)
    sel = np.arange(0, n, 23)
    dist = orc.make_dist_periodic1d(float(n))
    ref = []
    for j in sel:
        r = np.eye(k)
        for _ in range(n_iter):
            r = orc.lienks_weights_point(data["grid_rows"][j], r, perts, data["normed_obs"][None], data["obs_rows"], dist, (20.,),
                                         tau, eps)
        ref.append(r)
    ref = np.stack(ref)
    _close(w.cpu().numpy()[sel], ref, TOL9)
    _close(xa.cpu().numpy().reshape(1, 1, k, n)[..., sel], orc.apply_weights(data["state"][..., sel], ref),
                               TOL9)


def test_multi_chunk_identities_sphere():
    """More grid points than one Gram-scratch chunk holds (the scratch is capped at 2 GiB = about 150 000 slots at k = 50, so
    302 500 points run as three chunks): the IEnKS first iteration and the linear kernel program must reproduce the LETKF
    weights in every chunk (slot offsets, per-chunk flags and scratch reuse of k_ienks_pre / k_ienks_keep / k_kernelise)."""
    from pytassim_b200 import kernels as K
    k = 50
    data = syn.sphere_latlon(550, 550, k, 200_000, seed=5)
    n = data["state"].shape[-1]
    x = torch.as_tensor(data["state"].reshape(1, k, n)).cuda()
    metric = m.HaversineDistance(6371.0)

    def engine(kernel=None):
        eng = LETKFEngine(k, 1, metric, 500.0, inf_factor=1.0)
        if kernel is not None:
            eng.set_kernel(kernel)
        eng.set_grid(data["grid_rows"][:, 1:])
        eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
        return eng
    eng = engine()
    xa_ref, w_ref = eng.analyse(x, return_weights=True)
    xa, w = eng.ienks_step(x, np.eye(k), tau=1.0)
    scale = float(w_ref.abs().max())
    assert float((w - w_ref).abs().max()) <= 1e-10 * scale
    assert float((xa - xa_ref).abs().max()) <= 1e-10 * float(xa_ref.abs().max())
    del w, xa
    xa2, w2 = engine(K.LinearKernel() + K.ScaleKernel(0.)).analyse(x, return_weights=True)
    assert float((w2 - w_ref).abs().max()) <= 1e-10 * scale
    assert float((xa2 - xa_ref).abs().max()) <= 1e-10 * float(xa_ref.abs().max())
