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
