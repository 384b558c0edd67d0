"""Generate the golden vectors under tests/golden/ from the reference's OWN code.

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs /root/reference, which does not
exist on the GPU box).  It loads the reference's numerical leaf modules *unchanged, by file path*:

    pytassim/core/base.py, core/utils.py, core/etkf.py          (torch)
    pytassim/localization/localization.py, gaspari_cohn.py      (numpy)
    pytassim/interface/wrapper.py                                (torch + numpy)

under a stub ``pytassim`` package (the real package cannot be imported: xarray / dask / netCDF4 are
not installed) with one shim: ``torch.symeig`` (removed from torch; called at core/utils.py:57) is
mapped to ``torch.linalg.eigh``.  The reference fixtures tests/data/test_state.nc and
test_single_obs.nc are netCDF-3 and are read with scipy.

Usage:  python oracle/make_golden.py            (writes tests/golden/*.npz)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PYTASSIM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def load_reference_leaves(ref=REF):
    """Import the reference leaves under a stub package; returns a namespace."""
    # torch >= 1.13 keeps a ``symeig`` stub that only raises; replace it unconditionally
    def _symeig(t, eigenvectors=True, upper=False):
        return torch.linalg.eigh(t, UPLO="U" if upper else "L")
    torch.symeig = _symeig

    def stub(name):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
        return mod

    for pkg in ("pytassim", "pytassim.core", "pytassim.localization", "pytassim.interface"):
        if pkg not in sys.modules:
            stub(pkg)

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    ns = types.SimpleNamespace()
    ns.core_base = load("pytassim.core.base", "pytassim/core/base.py")
    ns.core_utils = load("pytassim.core.utils", "pytassim/core/utils.py")
    ns.core_etkf = load("pytassim.core.etkf", "pytassim/core/etkf.py")
    ns.loc_base = load("pytassim.localization.localization", "pytassim/localization/localization.py")
    ns.loc_gc = load("pytassim.localization.gaspari_cohn", "pytassim/localization/gaspari_cohn.py")
    ns.wrapper = load("pytassim.interface.wrapper", "pytassim/interface/wrapper.py")
    return ns


def read_fixtures(ref=REF):
    """tests/data/test_state.nc (2,3,10,40) and test_single_obs.nc -> plain arrays."""
    from scipy.io import netcdf_file
    st = netcdf_file(os.path.join(ref, "tests/data/test_state.nc"), mmap=False)
    state = np.array(st.variables["__xarray_dataarray_variable__"].data, dtype=np.float64)
    grid = np.array(st.variables["grid"].data, dtype=np.float64)
    hours = np.array(st.variables["time"].data, dtype=np.float64)
    ob = netcdf_file(os.path.join(ref, "tests/data/test_single_obs.nc"), mmap=False)
    obs = np.array(ob.variables["observations"].data, dtype=np.float64)
    cov = np.array(ob.variables["covariance"].data, dtype=np.float64)
    obs_grid = np.array(ob.variables["obs_grid_1"].data, dtype=np.float64)
    # 'hours since 1992-12-25 00:00:00' -> unix seconds (utilities/pandas.py:28-45)
    t0 = (np.datetime64("1992-12-25T00:00:00") - np.datetime64("1970-01-01T00:00:00")) / np.timedelta64(1, "s")
    t_unix = t0 + 3600.0 * hours
    return dict(state=state, grid=grid, t_unix=t_unix, obs=obs, cov=cov, obs_grid=obs_grid)


def reference_letkf(ref, state, normed_perts, normed_obs, grid_rows, obs_rows, localization, inf_factor,
                    grid_subset=None):
    """The reference's hot loop, with the reference's own functions: for every grid point
    ``wrapper_localization(wrapper_bridge(ETKFModule))`` (interface/letkf.py:127-143), then the
    numpy form of ``_apply_weights`` (interface/base.py:257-278)."""
    module = ref.core_etkf.ETKFModule(inf_factor=torch.tensor(inf_factor, dtype=torch.float64))
    bridged = ref.wrapper.wrapper_bridge(module, torch.device("cpu"), torch.float64)
    localized = ref.wrapper.wrapper_localization(bridged, localization)
    sel = np.arange(state.shape[-1]) if grid_subset is None else np.asarray(grid_subset)
    weights, lists, lws = [], [], []
    for g in sel:
        weights.append(localized(grid_rows[g], normed_perts, normed_obs, obs_info=obs_rows))
        if localization is not None:
            use, w = localization.localize_obs(grid_rows[g], obs_rows)
            lists.append(np.nonzero(use)[0].astype(np.int32))
            lws.append(w[use])
    weights = np.stack(weights, axis=0)
    mean = state[..., sel].mean(axis=2, keepdims=True)
    perts = state[..., sel] - mean
    analysis = mean + np.einsum("vtig,gij->vtjg", perts, weights)
    return analysis, weights, lists, lws


def load_reference_ketkf(ref=REF):
    """core/ketkf.py + kernels/{utils,base_kernels,linear}.py, unchanged, by file path (after load_reference_leaves)."""
    if "pytassim.kernels" not in sys.modules:
        mod = types.ModuleType("pytassim.kernels"); mod.__path__ = []; sys.modules["pytassim.kernels"] = mod

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    load("pytassim.kernels.utils", "pytassim/kernels/utils.py")
    load("pytassim.kernels.base_kernels", "pytassim/kernels/base_kernels.py")
    linear = load("pytassim.kernels.linear", "pytassim/kernels/linear.py")
    core = load("pytassim.core.ketkf", "pytassim/core/ketkf.py")
    return core, linear


def make_ketkf_golden():
    """tests/golden/ketkf_linear.npz: KETKFModule(LinearKernel) of the reference on seeded inputs (centred and uncentred
    perturbations, empty observations) and the localized KETKF on the reference fixtures (interface/lketkf.py path:
    wrapper_localization(wrapper_bridge(KETKFModule)) per grid point, GaspariCohn((10.,), |grid - obs|))."""
    ref = load_reference_leaves()
    core, linear = load_reference_ketkf()
    out = {}
    rng = np.random.RandomState(1234)
    for i, (k, p, rho) in enumerate([(2, 1, 1.0), (5, 3, 1.0), (10, 40, 1.1), (40, 38, 1.2), (50, 200, 1.05)]):
        hx = rng.normal(size=(k, p))
        perts = hx - hx.mean(axis=0, keepdims=True)
        obs = rng.normal(size=(1, p))
        module = core.KETKFModule(kernel=linear.LinearKernel(), inf_factor=torch.tensor(rho, dtype=torch.float64))
        out["c{0}_perts".format(i)] = perts; out["c{0}_raw".format(i)] = hx; out["c{0}_obs".format(i)] = obs
        out["c{0}_rho".format(i)] = np.float64(rho)
        out["c{0}_w".format(i)] = module(torch.as_tensor(perts), torch.as_tensor(obs)).numpy()
        out["c{0}_w_raw".format(i)] = module(torch.as_tensor(hx), torch.as_tensor(obs)).numpy()
    out["n_cases"] = np.int64(5)
    module = core.KETKFModule(kernel=linear.LinearKernel(), inf_factor=torch.tensor(1.3, dtype=torch.float64))
    out["empty_w"] = module(torch.zeros((6, 0), dtype=torch.float64), torch.zeros((1, 0), dtype=torch.float64)).numpy()
    # localized KETKF on the fixtures, first time slice (the setup of interface/test_lketkf.py / test_letkf.py:106-157)
    fx = read_fixtures()
    state = fx["state"][:, :1]
    hx = state[0, 0]                                            # dummy obs operator: variable 'x' (testing/dummy.py:39-66)
    mean = hx.mean(axis=0)
    rc = 1.0 / np.sqrt(np.diag(fx["cov"]))
    perts = (hx - mean) * rc
    innov = (fx["obs"][0] - mean) * rc
    grid_rows = np.stack([np.full(40, fx["t_unix"][0]), fx["grid"]], axis=1)
    obs_rows = np.stack([np.full(40, fx["t_unix"][0]), fx["obs_grid"]], axis=1)
    loc = ref.loc_gc.GaspariCohn((10.,), lambda g, o: np.abs(g[1] - np.asarray(o)[:, 1]))
    module = core.KETKFModule(kernel=linear.LinearKernel(), inf_factor=torch.tensor(1.1, dtype=torch.float64))
    bridged = ref.wrapper.wrapper_bridge(module, torch.device("cpu"), torch.float64)
    localized = ref.wrapper.wrapper_localization(bridged, loc)
    weights = np.stack([localized(grid_rows[g], perts, innov[None], obs_info=obs_rows) for g in range(40)])
    smean = state.mean(axis=2, keepdims=True)
    analysis = smean + np.einsum('vtig,gij->vtjg', state - smean, weights)          # interface/base.py:257-278
    out.update(lketkf_state=state, lketkf_perts=perts, lketkf_innov=innov, lketkf_weights=weights, lketkf_analysis=analysis,
               lketkf_grid=fx["grid"], lketkf_obs_grid=fx["obs_grid"], lketkf_obs=fx["obs"][:1], lketkf_cov=fx["cov"])
    np.savez(os.path.join(OUT, "ketkf_linear.npz"), **out)
    print("wrote ketkf_linear.npz")


def load_reference_kernels(ref=REF):
    """Every module of pytassim/kernels the device path covers, unchanged, by file path -> namespace of the classes."""
    load_reference_leaves()
    core, linear = load_reference_ketkf()

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ref, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    ns = types.SimpleNamespace(LinearKernel=linear.LinearKernel)
    for modname, classes in (("rbf", ("GaussKernel", "RBFKernel")), ("polynomial", ("PolyKernel",)), ("tanh", ("TanhKernel",)),
                             ("rational", ("RationalKernel",)), ("scale", ("ScaleKernel",)), ("diag", ("DiagKernel",)),
                             ("orn_uhl", ("OrnsteinUhlenbeckKernel",))):
        mod = load("pytassim.kernels." + modname, "pytassim/kernels/{0}.py".format(modname))
        for cls in classes:
            setattr(ns, cls, getattr(mod, cls))
    return core, ns


def make_kernels_golden():
    """tests/golden/ketkf_kernels.npz: ``KETKFModule(kernel)`` of the reference (core/ketkf.py + kernels/*.py) for every
    configuration of pytassim_b200/testing/kernel_cases.py on seeded inputs, and the localized KETKF
    (wrapper_localization(wrapper_bridge(KETKFModule)), interface/lketkf.py:84-110) on the reference fixtures."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "torch-assimilate_b200", "pytassim_b200", "testing"))
    import kernel_cases as kc
    ref = load_reference_leaves()
    core, ns = load_reference_kernels()
    out = {}
    rng = np.random.RandomState(4321)
    for i, (k, p, rho) in enumerate(kc.PROBLEM_SIZES):
        hx = rng.normal(size=(k, p))
        perts = hx - hx.mean(axis=0, keepdims=True)
        obs = rng.normal(size=(1, p))
        out["c{0}_perts".format(i)] = perts; out["c{0}_obs".format(i)] = obs; out["c{0}_rho".format(i)] = np.float64(rho)
        for name, build in kc.KERNEL_CASES:
            module = core.KETKFModule(kernel=build(ns, p), inf_factor=torch.tensor(rho, dtype=torch.float64))
            out["c{0}_w_{1}".format(i, name)] = module(torch.as_tensor(perts), torch.as_tensor(obs)).numpy()
    out["n_cases"] = np.int64(len(kc.PROBLEM_SIZES))
    # the L1 kernel the device rejects: pins the oracle's restatement only
    module = core.KETKFModule(kernel=ns.OrnsteinUhlenbeckKernel(lengthscale=30.), inf_factor=torch.tensor(1.1, dtype=torch.float64))
    out["c0_w_ornuhl"] = module(torch.as_tensor(out["c0_perts"]), torch.as_tensor(out["c0_obs"])).numpy()
    fx = read_fixtures()
    state = fx["state"][:, :1]
    hx = state[0, 0]                                            # dummy obs operator: variable 'x' (testing/dummy.py:39-66)
    mean = hx.mean(axis=0)
    rc = 1.0 / np.sqrt(np.diag(fx["cov"]))
    perts = (hx - mean) * rc
    innov = (fx["obs"][0] - mean) * rc
    grid_rows = np.stack([np.full(40, fx["t_unix"][0]), fx["grid"]], axis=1)
    obs_rows = np.stack([np.full(40, fx["t_unix"][0]), fx["obs_grid"]], axis=1)
    loc = ref.loc_gc.GaspariCohn((10.,), lambda g, o: np.abs(g[1] - np.asarray(o)[:, 1]))
    smean = state.mean(axis=2, keepdims=True)
    for name, build in kc.KERNEL_CASES:
        module = core.KETKFModule(kernel=build(ns, 20), inf_factor=torch.tensor(1.1, dtype=torch.float64))
        bridged = ref.wrapper.wrapper_bridge(module, torch.device("cpu"), torch.float64)
        localized = ref.wrapper.wrapper_localization(bridged, loc)
        weights = np.stack([localized(grid_rows[g], perts, innov[None], obs_info=obs_rows) for g in range(40)])
        out["lketkf_weights_" + name] = weights
        out["lketkf_analysis_" + name] = smean + np.einsum('vtig,gij->vtjg', state - smean, weights)   # interface/base.py:257-278
    out.update(lketkf_state=state, lketkf_perts=perts, lketkf_innov=innov, lketkf_grid=fx["grid"],
               lketkf_obs_grid=fx["obs_grid"])
    np.savez_compressed(os.path.join(OUT, "ketkf_kernels.npz"), **out)
    print("wrote ketkf_kernels.npz", os.path.getsize(os.path.join(OUT, "ketkf_kernels.npz")))


def make_ienks_golden():
    """tests/golden/ienks.npz: ``IEnKSTransformModule`` / ``IEnKSBundleModule`` of the reference (core/ienks.py, unchanged, by
    file path) iterated from the prior identity weights on seeded inputs, and the localized IEnKS
    (wrapper_localization(wrapper_bridge(module)) with args_to_skip=(0,), interface/lienks.py:68-118) on the reference
    fixtures for three iterations.  The bundle inputs are scaled by epsilon, as the ensemble the reference propagates for the
    bundle variant is (interface/ienks.py:153-160)."""
    ref = load_reference_leaves()
    spec = importlib.util.spec_from_file_location("pytassim.core.ienks", os.path.join(REF, "pytassim/core/ienks.py"))
    ienks = importlib.util.module_from_spec(spec)
    sys.modules["pytassim.core.ienks"] = ienks
    spec.loader.exec_module(ienks)
    t64 = lambda v: torch.tensor(v, dtype=torch.float64)
    out = {}
    rng = np.random.RandomState(777)
    cases = [(5, 3, 1.0), (10, 40, 1.0), (24, 60, 0.5), (40, 38, 0.8), (50, 200, 1.0)]
    for i, (k, p, tau) in enumerate(cases):
        hx = rng.normal(size=(k, p))
        perts = hx - hx.mean(axis=0, keepdims=True)
        obs = rng.normal(size=(1, p))
        out["c{0}_perts".format(i)] = perts; out["c{0}_obs".format(i)] = obs; out["c{0}_tau".format(i)] = np.float64(tau)
        for variant, eps in (("transform", None), ("bundle", 1e-2)):
            module = ienks.IEnKSTransformModule(tau=t64(tau)) if eps is None else ienks.IEnKSBundleModule(epsilon=t64(eps), tau=t64(tau))
            scale = 1.0 if eps is None else eps
            w = torch.eye(k, dtype=torch.float64)
            for it in range(3):                             # same observation-space variables every iteration: pins the core only
                w = module(w, torch.as_tensor(perts * scale), torch.as_tensor(obs))
                out["c{0}_{1}_w{2}".format(i, variant, it)] = w.numpy().copy()
        out["c{0}_eps".format(i)] = np.float64(1e-2)
    out["n_cases"] = np.int64(len(cases))
    module = ienks.IEnKSTransformModule(tau=t64(0.7))
    w_in = rng.normal(size=(6, 6))
    out["empty_in"] = w_in
    out["empty_w"] = module(torch.as_tensor(w_in), torch.zeros((6, 0), dtype=torch.float64), torch.zeros((1, 0), dtype=torch.float64)).numpy()
    # localized IEnKS on the fixtures, first time slice (setup of interface/test_lienks.py = test_letkf.py:106-157)
    fx = read_fixtures()
    state = fx["state"][:, :1]
    hx = state[0, 0]
    mean = hx.mean(axis=0)
    rc = 1.0 / np.sqrt(np.diag(fx["cov"]))
    perts = (hx - mean) * rc
    innov = (fx["obs"][0] - mean) * rc
    grid_rows = np.stack([np.full(40, fx["t_unix"][0]), fx["grid"]], axis=1)
    obs_rows = np.stack([np.full(40, fx["t_unix"][0]), fx["obs_grid"]], axis=1)
    loc = ref.loc_gc.GaspariCohn((10.,), lambda g, o: np.abs(g[1] - np.asarray(o)[:, 1]))
    smean = state.mean(axis=2, keepdims=True)
    for variant, eps, tau in (("transform", None, 0.6), ("bundle", 1e-2, 1.0)):
        module = ienks.IEnKSTransformModule(tau=t64(tau)) if eps is None else ienks.IEnKSBundleModule(epsilon=t64(eps), tau=t64(tau))
        bridged = ref.wrapper.wrapper_bridge(module, torch.device("cpu"), torch.float64)
        localized = ref.wrapper.wrapper_localization(bridged, loc)
        scale = 1.0 if eps is None else eps
        weights = np.stack([np.eye(10)] * 40)
        for it in range(3):
            weights = np.stack([localized(grid_rows[g], weights[g], perts * scale, innov[None], obs_info=obs_rows,
                                          args_to_skip=(0, )) for g in range(40)])
            out["l_{0}_w{1}".format(variant, it)] = weights.copy()
        out["l_{0}_analysis".format(variant)] = smean + np.einsum('vtig,gij->vtjg', state - smean, weights)
        out["l_{0}_tau".format(variant)] = np.float64(tau)
    out.update(l_state=state, l_perts=perts, l_innov=innov, l_grid=fx["grid"], l_obs_grid=fx["obs_grid"], l_eps=np.float64(1e-2))
    np.savez_compressed(os.path.join(OUT, "ienks.npz"), **out)
    print("wrote ienks.npz", os.path.getsize(os.path.join(OUT, "ienks.npz")))


def make_product_golden():
    """tests/golden/product_loc.npz: the reference's GaspariCohn with a dist_func that returns TWO rows (horizontal ring
    distance x |level difference|) and two length scales: localize_obs for a few grid rows, and the LETKF analysis of the
    reference's hot loop on seeded inputs (gaspari_cohn.py:124-135, interface/letkf.py:127-143)."""
    ref = load_reference_leaves()
    rng = np.random.RandomState(77)
    n_grid, k, n_lev = 48, 12, 4
    pos = np.tile(np.arange(n_grid // n_lev, dtype=np.float64), n_lev)
    lev = np.repeat(np.arange(n_lev, dtype=np.float64), n_grid // n_lev)
    grid_rows = np.stack([np.zeros(n_grid), pos, lev], axis=1)
    m = 60
    obs_rows = np.stack([np.zeros(m), rng.uniform(0, n_grid // n_lev, size=m), rng.randint(0, n_lev, size=m).astype(np.float64)
                         + rng.uniform(-0.3, 0.3, size=m)], axis=1)
    period = float(n_grid // n_lev)

    def dist(g, o):
        o = np.asarray(o)
        d = np.abs(g[1] - o[:, 1])
        return np.stack([np.minimum(d, period - d), np.abs(g[2] - o[:, 2])], axis=0)
    loc = ref.loc_gc.GaspariCohn((3.0, 1.5), dist)
    use = np.stack([loc.localize_obs(grid_rows[g], obs_rows)[0] for g in range(n_grid)])
    w = np.stack([loc.localize_obs(grid_rows[g], obs_rows)[1] for g in range(n_grid)])
    state = rng.normal(size=(1, 1, k, n_grid))
    hx = rng.normal(size=(k, m))
    perts = hx - hx.mean(axis=0, keepdims=True)
    innov = rng.normal(size=m)
    ana, weights, lists, lws = reference_letkf(ref, state, perts, innov[None], grid_rows, obs_rows, loc, 1.1)
    off, idx = csr(lists)
    np.savez(os.path.join(OUT, "product_loc.npz"), grid_rows=grid_rows, obs_rows=obs_rows, period=np.float64(period),
             radius=np.array([3.0, 1.5]), use=use, w=w, state=state, perts=perts, innov=innov, analysis=ana, weights=weights,
             csr_off=off, csr_idx=idx, rho=np.float64(1.1))
    print("wrote product_loc.npz")


def csr(lists):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in lists])
    idx = np.concatenate(lists) if lists else np.zeros(0, dtype=np.int32)
    return off, idx.astype(np.int32)


def main():
    sys.path.insert(0, HERE)
    import letkf_oracle as orc
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference_leaves()
    fx = read_fixtures()

    # ---- 1. core KAT from tests/unit_tests/core/test_etkf.py:47-103 (2 members, 1 obs) -------------
    ens_obs = np.array([0.5, -0.5]); obs = np.array([0.2]); obs_var = np.array([[0.5]])
    cinv = np.linalg.inv(np.linalg.cholesky(obs_var))
    normed_perts = (ens_obs.reshape(2, 1) @ cinv)
    normed_obs = ((obs - ens_obs.mean()) @ cinv).reshape(1, 1)
    mod = ref.core_etkf.ETKFModule()
    w_mean, w_perts, cov = mod._estimate_weights(torch.from_numpy(normed_perts), torch.from_numpy(normed_obs))
    W = mod(torch.from_numpy(normed_perts), torch.from_numpy(normed_obs))
    np.savez(os.path.join(OUT, "core_kat.npz"), normed_perts=normed_perts, normed_obs=normed_obs,
             w_mean=w_mean.numpy(), w_perts=w_perts.numpy(), cov=cov.numpy(), W=W.numpy())

    # ---- 2. core on random inputs, several (k, p, rho) ---------------------------------------------
    rnd = np.random.RandomState(42)
    cases = {}
    for n, (k, p, rho) in enumerate([(10, 40, 1.0), (50, 19, 1.1), (40, 38, 1.1), (16, 300, 1.3),
                                     (50, 1, 1.0), (3, 7, 0.9), (64, 200, 1.1), (100, 500, 1.05)]):
        Y = rnd.normal(size=(k, p)); Y -= Y.mean(axis=0, keepdims=True)
        d = rnd.normal(size=(p,))
        m = ref.core_etkf.ETKFModule(inf_factor=torch.tensor(rho, dtype=torch.float64))
        cases[f"Y{n}"] = Y; cases[f"d{n}"] = d; cases[f"rho{n}"] = np.float64(rho)
        cases[f"W{n}"] = m(torch.from_numpy(Y), torch.from_numpy(d)).numpy()
    m = ref.core_etkf.ETKFModule(inf_factor=torch.tensor(1.1, dtype=torch.float64))
    cases["W_empty"] = m(torch.ones(10, 0, dtype=torch.float64), torch.ones(1, 0, dtype=torch.float64)).numpy()
    np.savez(os.path.join(OUT, "core_random.npz"), **cases)

    # ---- 3. Gaspari-Cohn tapers on the fixture grid (test_gaspari_cohn.py:52-171) ------------------
    grid = fx["grid"]
    dummy_distance = lambda a, b: np.abs(a - b)          # pytassim/testing/dummy.py:142-151
    gc = ref.loc_gc.GaspariCohn(5., dist_func=dummy_distance)
    gci = ref.loc_gc.GaspariCohnInf(5., dist_func=dummy_distance)
    dist = dummy_distance(10, grid)
    r = np.concatenate([np.linspace(0, 2.5, 4001), rnd.uniform(0, 2.2, 4000), [1.0, 2.0, 0.5, 1.5, 0.0]])
    with np.errstate(all="ignore"):
        out = dict(dist=dist, r=r,
                   gc_f1=gc._f1(dist), gc_f2=gc._f2(dist), gci_f1=gci._f1(dist), gci_f2=gci._f2(dist),
                   gci_f3=gci._f3(dist), gci_f4=gci._f4(dist))
    for name, loc in (("gc", gc), ("gci", gci)):
        for g in (0, 10, 9999999):
            use, w = loc.localize_obs(g, grid)
            out[f"{name}_use_{g}"] = use; out[f"{name}_w_{g}"] = w
        loc_r = type(loc)(1., dist_func=lambda a, b: b)
        use, w = loc_r.localize_obs(0, r)
        out[f"{name}_use_r"] = use; out[f"{name}_w_r"] = w
    # two components with different radii (gaspari_cohn.py:126-134)
    gc2 = ref.loc_gc.GaspariCohn((5., 2.), dist_func=lambda a, b: (np.abs(a - b), np.abs(a - b) * 0.25))
    use, w = gc2.localize_obs(10, grid)
    out["gc2_use_10"] = use; out["gc2_w_10"] = w
    np.savez(os.path.join(OUT, "gaspari_cohn.npz"), **out)

    # ---- 4. fixtures through the whole path: tests/unit_tests/interface/test_letkf.py:106-157 ------
    state = fx["state"]; k = state.shape[2]
    # dummy_obs_operator (pytassim/testing/dummy.py:39-66): var 'x' -> (ensemble,time,obs_grid_1) view
    def hx_of(st):
        return np.transpose(st[0], (1, 0, 2))             # (k, n_t, n_grid)
    res = dict(state=state, obs=fx["obs"], cov=fx["cov"], grid=grid, t_unix=fx["t_unix"], obs_grid=fx["obs_grid"])
    # (a) time slice 0, single obs dataset, GC(10) with |grid - obs|
    st0 = state[:, :1]
    innov, perts = orc.obs_space_variables([hx_of(st0)], [fx["obs"][:1]], [fx["cov"]])
    grid_rows = np.stack([np.full(grid.shape, fx["t_unix"][0]), grid], axis=1)
    obs_rows = np.stack([np.full(fx["obs_grid"].shape, fx["t_unix"][0]), fx["obs_grid"]], axis=1)
    res["a_innov"] = innov; res["a_perts"] = perts; res["a_grid_rows"] = grid_rows; res["a_obs_rows"] = obs_rows
    gc10 = ref.loc_gc.GaspariCohn((10.,), dist_func=lambda x, y: (np.abs(x[1] - y[:, 1]),))
    ana, W, lists, lws = reference_letkf(ref, st0, perts, innov, grid_rows, obs_rows, gc10, 1.0)
    res["a_analysis"] = ana; res["a_weights"] = W
    res["a_csr_off"], res["a_csr_idx"] = csr(lists); res["a_csr_w"] = np.concatenate(lws)
    # (b) last time slice, the obs dataset twice (test_letkf.py:64-70), no localization == global ETKF
    st2 = state[:, 2:]
    innov2, perts2 = orc.obs_space_variables([hx_of(st2)] * 2, [fx["obs"][2:]] * 2, [fx["cov"]] * 2)
    module = ref.core_etkf.ETKFModule(inf_factor=torch.tensor(1.0, dtype=torch.float64))
    bridged = ref.wrapper.wrapper_bridge(module, torch.device("cpu"), torch.float64)
    Wg = bridged(perts2, innov2)
    mean = st2.mean(axis=2, keepdims=True)
    res["b_innov"] = innov2; res["b_perts"] = perts2; res["b_weights"] = Wg
    res["b_analysis"] = mean + np.einsum("vtig,ij->vtjg", st2 - mean, Wg)
    # (c) GaspariCohnInf(8) and inflation 1.1 on slice 1
    st1 = state[:, 1:2]
    innov1, perts1 = orc.obs_space_variables([hx_of(st1)], [fx["obs"][1:2]], [fx["cov"]])
    # GaspariCohnInf sizes its weights by obs_grid.shape[-1] (gaspari_cohn.py:244), so it only works
    # when the obs info is laid out (ncol, M): hand it the transposed rows.
    gci8 = ref.loc_gc.GaspariCohnInf(8., dist_func=lambda x, y: np.abs(x[1] - y[1]))
    ana, W, lists, lws = reference_letkf(ref, st1, perts1, innov1, grid_rows, obs_rows.T, gci8, 1.1)
    res["c_innov"] = innov1; res["c_perts"] = perts1; res["c_analysis"] = ana; res["c_weights"] = W
    res["c_csr_off"], res["c_csr_idx"] = csr(lists)
    np.savez(os.path.join(OUT, "fixture_letkf.npz"), **res)

    # ---- 5. seeded synthetic shapes (BASELINE configs, scaled so the CPU reference finishes) ------
    # Inputs are NOT stored: tests regenerate them from the seed with the same generator
    # (torch-assimilate_b200/pytassim_b200/testing/synthetic.py); only reference outputs are stored.
    spec = importlib.util.spec_from_file_location(
        "synthetic", os.path.join(os.path.dirname(HERE), "torch-assimilate_b200", "pytassim_b200", "testing",
                                  "synthetic.py"))
    syn = importlib.util.module_from_spec(spec); spec.loader.exec_module(syn)

    def run(name, data, dist, radius, rho, sel, n_w=8, **extra):
        loc = ref.loc_gc.GaspariCohn(radius, dist_func=dist)
        ana, W, lists, lws = reference_letkf(ref, data["state"], data["normed_perts"], data["normed_obs"],
                                             data["grid_rows"], data["obs_rows"], loc, rho, grid_subset=sel)
        off, idx = csr(lists)
        np.savez(os.path.join(OUT, name), sel=sel, analysis=ana, weights=W[:n_w], csr_off=off, csr_idx=idx,
                 csr_w=np.concatenate(lws), radius=radius, rho=rho, **extra)

    # cfg1: Lorenz-96 N=40, k=50, all observed, periodic distance, GC c=5, rho=1.1
    data = syn.lorenz96_1d(40, 50, 1, seed=42)
    run("cfg1_l96_n40_k50.npz", data, orc.make_dist_periodic1d(40.0), 5.0, 1.1, np.arange(40), n_w=40,
        seed=42, n_grid=40, k=50, obs_stride=1)
    # cfg2 scaled to N=2000: k=40, every 2nd observed, periodic, c=20
    data = syn.lorenz96_1d(2000, 40, 2, seed=43)
    run("cfg2_l96_n2000_k40.npz", data, orc.make_dist_periodic1d(2000.0), 20.0, 1.1, np.arange(0, 2000, 37),
        seed=43, n_grid=2000, k=40, obs_stride=2)
    # benchmark_letkf.py defaults scaled: non-periodic |x - y|, 1 obs per 10 grid points, k=50, c=20
    data = syn.lorenz96_1d(1000, 50, 10, seed=45)
    run("bench_default_n1000_k50.npz", data, orc.dist_abs1d, 20.0, 1.1, np.arange(0, 1000, 23),
        seed=45, n_grid=1000, k=50, obs_stride=10)
    # cfg3 scaled: 24x48 lat-lon grid, k=50, 3000 obs uniform on the sphere, haversine, c=1000 km
    data = syn.sphere_latlon(24, 48, 50, 3000, seed=44)
    run("cfg3_sphere_small.npz", data, orc.make_dist_haversine(6371.0), 1000.0, 1.1, np.arange(0, 24 * 48, 29),
        seed=44, nlat=24, nlon=48, k=50, n_obs=3000)
    # 2-D Euclidean, k=16, 2 state slices
    r = np.random.RandomState(46)
    n_grid, M, k = 400, 600, 16
    gx, gy = [a.reshape(-1) for a in np.meshgrid(np.arange(20.0), np.arange(20.0), indexing="ij")]
    ox, oy = r.uniform(-1, 20, size=M), r.uniform(-1, 20, size=M)
    st = r.normal(size=(2, 1, k, n_grid))
    hx = r.normal(size=(k, M)); y = r.normal(size=M)
    perts, innov = syn.obs_space_from_hx(hx, y, var=r.uniform(0.5, 2.0, size=M))
    data = dict(state=st, normed_perts=perts, normed_obs=innov,
                grid_rows=np.stack([np.zeros(n_grid), gx, gy], axis=1), obs_rows=np.stack([np.zeros(M), ox, oy], axis=1))
    run("euclid2d_k16.npz", data, orc.dist_euclid, 3.0, 1.05, np.arange(0, 400, 7), state=st, normed_perts=perts,
        normed_obs=innov, grid_rows=data["grid_rows"], obs_rows=data["obs_rows"])
    print("golden vectors written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  ", f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "ketkf":
    make_ketkf_golden()
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "kernels":
    make_kernels_golden()
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "ienks":
    make_ienks_golden()
    sys.exit(0)

if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "product":
    make_product_golden()
    sys.exit(0)

if __name__ == "__main__":
    main()
