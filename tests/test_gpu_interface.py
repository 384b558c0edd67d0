"""The reference-facing API on the GPU: ``LETKF(...).assimilate(state, observations, pseudo_state, analysis_time)`` and
``ETKF(...).assimilate`` on the reference's own fixtures, mirroring tests/unit_tests/interface/test_letkf.py and
test_etkf.py of the reference; expected values come from the reference's code (tests/golden/fixture_letkf.npz)."""
import numpy as np
import pandas as pd
import pytest

from pytassim_b200 import xrlite
from pytassim_b200.interface import ETKF, LETKF
from pytassim_b200.localization import GaspariCohn, GaspariCohnInf, AbsDistance1D, EuclideanDistance
from test_host_logic import _fixture_objects

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-10, atol=1e-10)


def test_letkf_localized_right(golden):
    """test_letkf.py:106-157: GaspariCohn((10.,), |grid - obs|), first time slice."""
    g, state, obs = _fixture_objects(golden)
    alg = LETKF(localization=GaspariCohn((10.,), AbsDistance1D()))
    st0, ob0 = state.isel(time=[0]), obs.isel(time=[0])
    ana = alg.assimilate(st0, ob0)
    assert ana.dims == state.dims and ana.shape == (2, 1, 10, 40)
    np.testing.assert_allclose(ana.values, g["a_analysis"], **TOL)
    # the full state with analysis_time = first time selects the same slice (filter mode, filter.py:149-152)
    ana2 = alg.assimilate(state, obs, analysis_time="1992-12-25 00:00")
    np.testing.assert_allclose(ana2.values, g["a_analysis"], **TOL)
    assert list(ana2.indexes["time"]) == [pd.Timestamp("1992-12-25 00:00")]


def test_letkf_without_localization_equals_etkf(golden):
    """test_letkf.py:64-70: two copies of the observations, last time step."""
    g, state, obs = _fixture_objects(golden)
    etkf_ana = ETKF().assimilate(state, (obs, obs))
    letkf_ana = LETKF().assimilate(state, (obs, obs))
    np.testing.assert_allclose(etkf_ana.values, g["b_analysis"], **TOL)
    np.testing.assert_allclose(letkf_ana.values, etkf_ana.values, **TOL)
    assert list(etkf_ana.indexes["time"]) == [pd.Timestamp("1992-12-25 02:00")]


def test_letkf_gcinf_inflation_second_slice(golden):
    g, state, obs = _fixture_objects(golden)
    alg = LETKF(localization=GaspariCohnInf(8., AbsDistance1D()), inf_factor=1.1)
    ana = alg.assimilate(state, obs, None, "1992-12-25 01:00")
    np.testing.assert_allclose(ana.values, g["c_analysis"], **TOL)


def test_letkf_multiindex_grid_and_engine_reuse(golden):
    """test_letkf.py:79-92: a MultiIndex grid (grid_point, height); Euclidean distance over both levels."""
    g, state, obs = _fixture_objects(golden)
    mi = pd.MultiIndex.from_product((np.arange(40), [0, ]), names=['grid_point', 'height'])
    state_mi = xrlite.DataArray(state.values, dict(var_name=state.indexes["var_name"], time=state.indexes["time"],
                                                   ensemble=state.indexes["ensemble"], grid=mi), state.dims)
    obs_mi = xrlite.Dataset({
        "observations": xrlite.DataArray(obs["observations"].values, dict(time=state.indexes["time"], obs_grid_1=mi),
                                         ("time", "obs_grid_1")),
        "covariance": obs["covariance"]})
    obs_mi.obs.operator = obs.obs.operator
    alg = LETKF(localization=GaspariCohn(10., EuclideanDistance(2)))
    ana = alg.assimilate(state_mi, obs_mi, analysis_time="1992-12-25 00:00")
    np.testing.assert_allclose(ana.values, g["a_analysis"], **TOL)          # height is constant: same distances
    ana_b = alg.assimilate(state_mi, obs_mi, analysis_time="1992-12-25 00:00")   # cached engine + grid
    np.testing.assert_array_equal(ana_b.values, ana.values)


def test_smoother_mode_uses_all_times(golden):
    """filter.py:147-152: smoother=True keeps every time and stacks the observations of all times."""
    g, state, obs = _fixture_objects(golden)
    ana = ETKF(smoother=True).assimilate(state, obs)
    assert ana.shape == state.shape
    import letkf_oracle as orc
    hx = np.transpose(g["state"][0], (1, 0, 2))
    innov, perts = orc.obs_space_variables([hx], [g["obs"]], [g["cov"]])
    ref, _ = orc.etkf_analysis(g["state"], perts, innov, 1.0)
    np.testing.assert_allclose(ana.values, ref, **TOL)


def test_localize_obs_signature(golden):
    """localization/localization.py:53-80 via the GPU neighbour search (test_gaspari_cohn.py:78-94)."""
    import letkf_oracle as orc
    loc = GaspariCohn(5., AbsDistance1D())
    grid = np.stack([np.zeros(40), np.arange(40.0)], axis=1)
    use, w = loc.localize_obs(np.array([0.0, 10.0]), grid)
    ruse, rw = orc.gaspari_cohn_localize(np.abs(10.0 - grid[:, 1]), 5.0)
    np.testing.assert_array_equal(use, ruse)
    np.testing.assert_allclose(w[use], rw[ruse], rtol=1e-12, atol=5e-15)
    use, _ = loc.localize_obs(np.array([0.0, 9999999.0]), grid)
    assert not use.any()


def test_device_obs_prep_bit_exact_against_oracle():
    """b200da_obs_prep == BaseAssimilation._get_obs_space_variables with a diagonal R
    (interface/base.py:359-379, observation.py:241-245): FP64 bit-exact, FP32 to rounding."""
    import numpy as np
    import torch
    import letkf_oracle as orc
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200.localization.metrics import AbsDistance1D
    rng = np.random.RandomState(3)
    k, n_t, n_o = 23, 3, 1777
    hx = rng.normal(size=(k, n_t, n_o)) * 3.0 + 1.5
    y = rng.normal(size=(n_t, n_o))
    var = rng.uniform(0.1, 4.0, size=n_o)
    innov_ref, perts_ref = orc.obs_space_variables([hx], [y], [var])
    eng = LETKFEngine(k, 1, AbsDistance1D(), 1.0, dtype=torch.float64)
    yn, d = eng.obs_prep(hx.reshape(k, -1), y.reshape(-1), np.tile(var, n_t))
    np.testing.assert_array_equal(yn.cpu().numpy(), perts_ref)
    np.testing.assert_array_equal(d.cpu().numpy(), innov_ref)
    eng32 = LETKFEngine(k, 1, AbsDistance1D(), 1.0, dtype=torch.float32)
    yn32, d32 = eng32.obs_prep(hx.reshape(k, -1), y.reshape(-1), np.tile(var, n_t))
    assert yn32.dtype == torch.float32
    np.testing.assert_allclose(yn32.cpu().numpy(), perts_ref, rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(d32.cpu().numpy(), innov_ref, rtol=2e-5, atol=2e-5)
    with pytest.raises(ValueError):
        eng.obs_prep(hx.reshape(k, -1), y.reshape(-1)[:-1], np.tile(var, n_t))


def _wrap_host_only(op):
    """The same operator without ``device_index``: forces the host operator path (base.py:181-220)."""
    return lambda obs_ds, st: op(obs_ds, st)


@pytest.mark.parametrize("smoother", [False, True])
def test_device_obs_operator_gather_equals_host_operator(smoother):
    """column-selecting operators (obs_ops/lorenz_96/identity.py, examples/benchmark_letkf.py:90-104) are
    gathered from the device-resident state and fused with the prep; the analysis is bit-identical to calling the operators on
    the host and uploading HX (FP64), for a localized LETKF and the global ETKF, filter and smoother mode, two datasets."""
    from pytassim_b200.obs_ops import IdentityOperator, NearestGridOperator
    from test_host_logic import _operator_objects, _obs_for
    rng, t, state = _operator_objects(n_grid=200, k=12, n_time=3, seed=8)
    pts = np.sort(np.random.RandomState(2).choice(200, size=90, replace=False))
    op1 = IdentityOperator(obs_points=pts, len_grid=200)
    op2 = NearestGridOperator(len_grid=200, nr_obs=37)
    times = t if smoother else t[[2]]

    def datasets(device):
        r = np.random.RandomState(77)
        ds1 = _obs_for(r, times, pts); ds1.obs.operator = op1 if device else _wrap_host_only(op1)
        ds2 = _obs_for(r, times[-1:], op2.obs_grid); ds2.obs.operator = op2 if device else _wrap_host_only(op2)
        return ds1, ds2
    for alg in (LETKF(localization=GaspariCohn(6., AbsDistance1D()), inf_factor=1.05, smoother=smoother),
                ETKF(inf_factor=1.05, smoother=smoother)):
        ana_dev = alg.assimilate(state, datasets(True))
        ana_host = alg.assimilate(state, datasets(False))
        assert ana_dev.shape == ((2, 3, 12, 200) if smoother else (2, 1, 12, 200))
        np.testing.assert_array_equal(ana_dev.values, ana_host.values)
        assert np.abs(ana_dev.values - state.values[:, -ana_dev.shape[1]:]).max() > 1e-3      # something was assimilated
    # a separate pseudo state (assimilate(..., pseudo_state=)) is what the operators read (base.py:342-357)
    pseudo = state.copy(data=state.values + 0.1 * rng.normal(size=state.values.shape))
    alg = LETKF(localization=GaspariCohn(6., AbsDistance1D()), smoother=smoother)
    np.testing.assert_array_equal(alg.assimilate(state, datasets(True), pseudo).values,
                                  alg.assimilate(state, datasets(False), pseudo).values)


def test_device_obs_gather_prep_fp32_and_bounds():
    from pytassim_b200.engine import LETKFEngine
    import torch
    rng = np.random.RandomState(6)
    k, n, m = 9, 1000, 333
    xp = rng.normal(size=(2, 2, k, n)).astype(np.float32)
    g_pos = rng.randint(0, n, size=m)
    src = ((1 * 2 + 1) * k * n + g_pos).astype(np.int64)
    y = rng.normal(size=m).astype(np.float32); var = (0.5 + rng.rand(m)).astype(np.float32)
    eng = LETKFEngine(k, 4, AbsDistance1D(), 1.0, dtype=torch.float32)
    yn, d = eng.obs_gather_prep(xp, src, n, y, var)
    hx = xp[1, 1][:, g_pos]
    yn_ref, d_ref = eng.obs_prep(hx, y, var)
    assert torch.equal(yn, yn_ref) and torch.equal(d, d_ref)
    with pytest.raises(IndexError):
        eng.obs_gather_prep(xp, src + 2 * k * n, n, y, var)


def test_letkf_with_all_ones_localization_equals_etkf(golden):
    """test_letkf.py:79-104: ``GaspariCohn((1., 1.), dist_func=zeros)`` (two radii, one distance row: only the first radius is
    used, gaspari_cohn.py:126-127) makes every observation local with weight 1, so the LETKF equals the global ETKF at every
    grid point (rtol = atol = 1e-10), with a plain and with a MultiIndex grid, two observation datasets."""
    from pytassim_b200.localization import ZeroDistance
    g, state, obs = _fixture_objects(golden)
    etkf_ana = ETKF().assimilate(state, (obs, obs))
    alg = LETKF(localization=GaspariCohn((1., 1.), dist_func=ZeroDistance()), chunksize=10)
    np.testing.assert_allclose(alg.assimilate(state, (obs, obs)).values, etkf_ana.values, **TOL)
    mi = pd.MultiIndex.from_product((np.arange(40), [0, ]), names=['grid_point', 'height'])
    state_mi = xrlite.DataArray(state.values, dict(var_name=state.indexes["var_name"], time=state.indexes["time"],
                                                   ensemble=state.indexes["ensemble"], grid=mi), state.dims)
    np.testing.assert_allclose(alg.assimilate(state_mi, (obs, obs)).values, etkf_ana.values, **TOL)
    use, w = alg.localization.localize_obs(np.array([0., 3.]), np.column_stack([np.zeros(5), np.arange(5.)]))
    assert use.all() and np.array_equal(w, np.ones(5))


def test_kernelised_etkf_linear_kernel(golden):
    """interface/ketkf.py, interface/lketkf.py with the default LinearKernel against the reference's KETKFModule
    (tests/golden/ketkf_linear.npz), same constructor signatures; objects that are no kernel descriptors raise (no CPU
    fallback); the non-linear kernels are covered by tests/test_gpu_kernels.py."""
    from pytassim_b200.interface import KETKF, LKETKF
    from pytassim_b200.kernels import LinearKernel
    g = golden("ketkf_linear.npz")
    _, state, obs = _fixture_objects(golden)
    st0, ob0 = state.isel(time=[0]), obs.isel(time=[0])
    alg = LKETKF(localization=GaspariCohn((10.,), AbsDistance1D()), kernel=LinearKernel(), inf_factor=1.1, smoother=False,
                 gpu=False, pre_transform=None, post_transform=None, chunksize=10, weight_save_path=None, forward_model=None)
    ana = alg.assimilate(st0, ob0)
    np.testing.assert_allclose(ana.values, g["lketkf_analysis"], **TOL)
    assert str(alg).startswith("Localized KETKF(inf_factor=1.1") and repr(KETKF(inf_factor=2.0)) == "KETKF(2.0,Linear)"
    glob = KETKF(inf_factor=1.1).assimilate(state, (obs, obs))
    np.testing.assert_allclose(glob.values, ETKF(inf_factor=1.1).assimilate(state, (obs, obs)).values, rtol=0, atol=0)
    with pytest.raises(NotImplementedError):
        KETKF(kernel=object())
    with pytest.raises(NotImplementedError):
        alg.kernel = "rbf"


def test_weight_save_path_exports_and_applies_stored_weights(golden, tmp_path):
    """interface/filter.py:159-162: with ``weight_save_path`` the estimated weights are written as netCDF, loaded back and
    applied; the analysis equals the fused path, and the stored weights equal the reference's (tests/golden)."""
    from pytassim_b200.utilities import load_netcdf
    g, state, obs = _fixture_objects(golden)
    st0, ob0 = state.isel(time=[0]), obs.isel(time=[0])
    path = str(tmp_path / "letkf_weights.nc")
    alg = LETKF(localization=GaspariCohn((10.,), AbsDistance1D()), weight_save_path=path)
    ana = alg.assimilate(st0, ob0)
    np.testing.assert_allclose(ana.values, g["a_analysis"], **TOL)
    stored = alg.load_weights()
    assert stored.dims == ('grid', 'ensemble', 'ensemble_new') and stored.shape == (40, 10, 10)
    np.testing.assert_allclose(stored.values, g["a_weights"], **TOL)
    assert list(stored.indexes['grid']) == list(state.indexes['grid'])
    # MultiIndex grid: levels encoded in the file, decoded on load
    mi = pd.MultiIndex.from_product((np.arange(40), [0, ]), names=['grid_point', 'height'])
    state_mi = xrlite.DataArray(st0.values, dict(var_name=state.indexes["var_name"], time=st0.indexes["time"],
                                                 ensemble=state.indexes["ensemble"], grid=mi), state.dims)
    ana_mi = LETKF(localization=GaspariCohn((10.,), AbsDistance1D()), weight_save_path=path).assimilate(state_mi, ob0)
    np.testing.assert_allclose(ana_mi.values, g["a_analysis"], **TOL)
    assert load_netcdf(path, array=True).indexes['grid'].equals(mi)
    # global ETKF
    gpath = str(tmp_path / "etkf_weights.nc")
    etkf_ana = ETKF(weight_save_path=gpath).assimilate(state, (obs, obs))
    np.testing.assert_allclose(etkf_ana.values, g["b_analysis"], **TOL)
    assert load_netcdf(gpath, array=True).dims == ('ensemble', 'ensemble_new')
