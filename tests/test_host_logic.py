"""CPU tests of the host side: C-ABI surface, interface validation / error behaviour, the xarray-glue restatement
against the golden vectors, metric objects against the oracle's distance functions."""
import ctypes
import os
import re
import warnings

import numpy as np
import pandas as pd
import pytest
import torch

import letkf_oracle as orc
from pytassim_b200 import _cabi, xrlite
from pytassim_b200.interface import ETKF, LETKF, StateError, ObservationError
from pytassim_b200.interface.base import BaseAssimilation, index_to_array, dtindex_to_total_seconds
from pytassim_b200.localization import (GaspariCohn, GaspariCohnInf, AbsDistance1D, PeriodicDistance1D,
                                        EuclideanDistance, HaversineDistance)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- C ABI -----------------------------------------------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b200da.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200da_[a-z0-9_]+)\s*\(", text)))


def test_cabi_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 25
    for name in syms:
        assert hasattr(lib, name), "libb200da.so does not export " + name


def test_cabi_binding_covers_header():
    assert sorted(_cabi.SIGNATURES) == _header_symbols()
    lib = _cabi.load()
    assert lib.b200da_version() >= 100
    assert lib.b200da_strerror(_cabi.ERR_SIZE).decode().startswith("observational size")
    assert lib.b200da_launch_count() >= 0


def test_plan_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _cabi.load()
    handle = ctypes.c_void_p()
    radius = (ctypes.c_double * 1)(1.0)
    rc = lib.b200da_plan_create(ctypes.byref(handle), 10, 1, 1, _cabi.METRIC_ABS1D, None, 0, radius, 1, 1e-5, 1.0,
                                _cabi.F64, _cabi.TAPER_GC)
    assert rc == _cabi.ERR_NO_DEVICE and not handle.value
    with pytest.raises(RuntimeError):
        from pytassim_b200.engine import LETKFEngine
        LETKFEngine(10, 1, AbsDistance1D(), 1.0)


def test_status_to_exception_mapping():
    with pytest.raises(ValueError):
        _cabi.check(_cabi.ERR_SIZE)
    with pytest.raises(NotImplementedError):
        _cabi.check(_cabi.ERR_UNSUPPORTED)
    with pytest.raises(_cabi.B200DAError):
        _cabi.check(_cabi.ERR_NO_DEVICE)
    _cabi.check(_cabi.OK)


# ---- metrics ---------------------------------------------------------------------------------------------------------
def test_metric_objects_equal_oracle_distance_functions():
    rnd = np.random.RandomState(3)
    obs = np.concatenate([np.zeros((200, 1)), rnd.uniform(0, 40, size=(200, 3))], axis=1)
    g = np.array([0.0, 7.25, 3.5, 1.0])
    np.testing.assert_array_equal(AbsDistance1D()(g, obs), orc.dist_abs1d(g, obs))
    np.testing.assert_array_equal(PeriodicDistance1D(40.0)(g, obs), orc.make_dist_periodic1d(40.0)(g, obs))
    np.testing.assert_array_equal(EuclideanDistance(3)(g, obs), orc.dist_euclid(g, obs))
    np.testing.assert_array_equal(EuclideanDistance(2)(g, obs), orc.dist_euclid(g[:3], obs[:, :3]))
    lat = np.degrees(np.arcsin(rnd.uniform(-1, 1, 200))); lon = rnd.uniform(0, 360, 200)
    obs = np.stack([np.zeros(200), lat, lon], axis=1)
    g = np.array([0.0, -33.3, 151.2])
    np.testing.assert_array_equal(HaversineDistance(6371.0)(g, obs), orc.make_dist_haversine(6371.0)(g, obs))
    # a pandas frame works like the reference's obs_info (interface/mixin_local.py:45-47)
    frame = pd.DataFrame(obs, columns=["time", "lat", "lon"])
    np.testing.assert_array_equal(HaversineDistance(6371.0)(g, frame), orc.make_dist_haversine(6371.0)(g, obs))


def test_foreign_dist_func_is_rejected():
    with pytest.raises(NotImplementedError):
        GaspariCohn(10.0, dist_func=lambda x, y: np.abs(x - y))
    with pytest.raises(NotImplementedError):
        GaspariCohnInf(10.0, dist_func=np.subtract)
    loc = GaspariCohn((10.0, 2.0), AbsDistance1D())
    assert str(loc) == 'GaspariCohn(l=[10.  2.])' and repr(loc) == 'GaspariCohn'


# ---- interface ---------------------------------------------------------------------------------------------------------
def _fixture_objects(golden, times=slice(None)):
    g = golden("fixture_letkf.npz")
    t = pd.to_datetime("1992-12-25") + pd.to_timedelta(np.arange(3), unit="h")
    state = xrlite.DataArray(g["state"], dict(var_name=["x", "y"], time=t, ensemble=np.arange(10), grid=np.arange(40)),
                             ("var_name", "time", "ensemble", "grid"))
    obs = xrlite.Dataset({
        "observations": xrlite.DataArray(g["obs"], dict(time=t, obs_grid_1=np.arange(40)), ("time", "obs_grid_1")),
        "covariance": xrlite.DataArray(g["cov"], dict(obs_grid_1=np.arange(40), obs_grid_2=np.arange(40)),
                                       ("obs_grid_1", "obs_grid_2")),
    })

    def dummy_obs_operator(obs_ds, st):                     # pytassim/testing/dummy.py:39-66
        x = st.isel(var_name=[0])
        return xrlite.DataArray(x.values[0], dict(time=obs_ds["observations"].indexes["time"], ensemble=st.indexes["ensemble"],
                                                  obs_grid_1=obs_ds["observations"].indexes["obs_grid_1"]),
                                ("time", "ensemble", "obs_grid_1"))
    obs.obs.operator = dummy_obs_operator
    return g, state, obs


def test_constructor_signatures_and_properties():
    alg = LETKF(localization=None, inf_factor=1.1, smoother=False, gpu=True, pre_transform=None, post_transform=None,
                chunksize=10, weight_save_path=None, forward_model=None)
    assert alg.chunks == {"grid": 10} and alg.dtype == torch.float64 and alg.device.type == "cuda"
    assert str(alg).startswith("Localized ETKF(inf_factor=1.1") and repr(ETKF(1.5)) == "ETKF(1.5)"
    with pytest.raises(TypeError):
        alg.dtype = float                                  # interface/base.py:115-118
    alg.inf_factor = 1.3
    assert abs(float(alg.inf_factor) - 1.3) < 1e-12
    with pytest.raises(NotImplementedError):
        ETKF(weight_save_path="weights.nc")


def test_validation_errors_and_warnings(golden):
    g, state, obs = _fixture_objects(golden)
    alg = LETKF()
    with pytest.warns(UserWarning):                        # interface/base.py:478-481
        assert alg.assimilate(state, []) is state
    with pytest.raises(TypeError):
        alg.assimilate(np.zeros((1, 1, 2, 2)), obs)
    bad = xrlite.DataArray(g["state"], {}, ("var", "time", "ensemble", "grid"))
    with pytest.raises(StateError):
        alg.assimilate(bad, obs)
    with pytest.raises(TypeError):
        alg.assimilate(state, (np.zeros(3),))
    broken = xrlite.Dataset({"observations": obs["observations"],
                             "covariance": xrlite.DataArray(np.ones(39), {}, ("obs_grid_1",))})
    with pytest.raises(ObservationError):
        alg.assimilate(state, broken)


def test_analysis_time_selection(golden):
    _, state, _ = _fixture_objects(golden)
    assert BaseAssimilation._get_analysis_time(state) == pd.Timestamp("1992-12-25 02:00")
    assert BaseAssimilation._get_analysis_time(state, "1992-12-25 01:00") == pd.Timestamp("1992-12-25 01:00")
    with pytest.warns(UserWarning):                        # interface/base.py:167-173
        assert BaseAssimilation._get_analysis_time(state, "1992-12-25 01:20") == pd.Timestamp("1992-12-25 01:00")


def test_obs_space_variables_match_reference_glue(golden):
    """interface/base.py:359-379 + :223-241 restated in the interface == the golden obs-space variables."""
    g, state, obs = _fixture_objects(golden)
    st0, ob0 = state.isel(time=[0]), obs.isel(time=[0])
    innov, perts, info = BaseAssimilation._get_obs_space_variables([ob0.obs.operator(ob0, st0)], [ob0])
    np.testing.assert_array_equal(innov, g["a_innov"]); np.testing.assert_array_equal(perts, g["a_perts"])
    np.testing.assert_array_equal(info, g["a_obs_rows"])
    st2, ob2 = state.isel(time=[2]), obs.isel(time=[2])
    hx = ob2.obs.operator(ob2, st2)
    innov, perts, info = BaseAssimilation._get_obs_space_variables([hx, hx], [ob2, ob2])
    np.testing.assert_array_equal(innov, g["b_innov"]); np.testing.assert_array_equal(perts, g["b_perts"])
    assert info.shape == (80, 2)
    # variance vector == diagonal covariance matrix (observation.py:241-275)
    diag = xrlite.Dataset({"observations": ob0["observations"],
                           "covariance": xrlite.DataArray(np.full(40, 0.5), {}, ("obs_grid_1",))})
    i2, p2, _ = BaseAssimilation._get_obs_space_variables([ob0.obs.operator(ob0, st0)], [diag])
    np.testing.assert_allclose(i2, g["a_innov"], rtol=1e-14); np.testing.assert_allclose(p2, g["a_perts"], rtol=1e-14)


def test_index_helpers():
    mi = pd.MultiIndex.from_product((np.arange(3), [0.5]), names=["grid_point", "height"])
    np.testing.assert_array_equal(index_to_array(mi), [[0, .5], [1, .5], [2, .5]])
    np.testing.assert_array_equal(index_to_array(pd.Index([3, 4])), [[3.], [4.]])
    t = pd.to_datetime(["1970-01-01 00:00:10", "1992-12-25 08:00:00"])
    np.testing.assert_array_equal(dtindex_to_total_seconds(t), [10.0, 725270400.0])


def test_datasets_without_operator_are_dropped(golden):
    """interface/base.py:213-218."""
    _, state, obs = _fixture_objects(golden)
    no_op = xrlite.Dataset(obs.data_vars)
    hx, kept = BaseAssimilation._apply_obs_operator(state, [no_op, obs])
    assert len(hx) == 1 and kept == [obs]
