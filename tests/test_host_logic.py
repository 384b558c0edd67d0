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


def test_cabi_constants_match_header_enums():
    """The Python constants of the binding are the enum values of include/b200da.h (guards against ABI drift)."""
    import re
    text = open(os.path.join(ROOT, "include", "b200da.h")).read()
    enums = {name: int(val) for name, val in re.findall(r"\b(B200DA_[A-Z0-9_]+)\s*=\s*(-?\d+)", text)}
    for name, val in enums.items():
        short = name[len("B200DA_"):]
        if hasattr(_cabi, short):
            assert getattr(_cabi, short) == val, name
    for short in ("KOP_LINEAR", "KOP_GAUSS", "KOP_POLY", "KOP_TANH", "KOP_RATIONAL", "KOP_SCALE", "KOP_DIAG", "KOP_ADD", "KOP_MUL",
                  "KOP_POW", "METRIC_HAVERSINE", "TAPER_GCINF", "F32", "SOLVER_JACOBI", "ERR_UNSUPPORTED"):
        assert "B200DA_" + short in enums and enums["B200DA_" + short] == getattr(_cabi, short), short
    assert int(re.search(r"#define B200DA_MAX_KERNEL_OPS (\d+)", text).group(1)) == _cabi.MAX_KERNEL_OPS


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
    assert ETKF(weight_save_path="weights.nc").weight_save_path == "weights.nc"      # interface/base.py:69


def test_weight_store_netcdf_roundtrip(tmp_path):
    """utilities/xarray.py:36-173 on the weight DataArray of interface/letkf.py:143-147: MultiIndex grid encoded as a range +
    ``multidim_levels`` + one coordinate per level, decoded on load; netCDF-3 layout of an unnamed xarray DataArray."""
    from scipy.io import netcdf_file
    from pytassim_b200.utilities import save_netcdf, load_netcdf, DATAARRAY_VARIABLE
    rng = np.random.RandomState(0)
    mi = pd.MultiIndex.from_product(([1.5, 2.5, 3.5], [0, 10]), names=['lat', 'lev'])
    w = xrlite.DataArray(rng.normal(size=(6, 4, 4)), dict(grid=mi, ensemble=np.arange(4), ensemble_new=np.arange(4)),
                         ('grid', 'ensemble', 'ensemble_new'))
    path = str(tmp_path / "weights.nc")
    assert save_netcdf(w, path) is None
    nc = netcdf_file(path, 'r', mmap=False)
    assert set(nc.variables) == {DATAARRAY_VARIABLE, 'grid', 'lat', 'lev', 'ensemble', 'ensemble_new'}
    assert nc.variables['grid']._attributes['multidim_levels'] == b'lat;lev'                     # xarray.py:84-89
    assert nc.variables[DATAARRAY_VARIABLE].dimensions == ('grid', 'ensemble', 'ensemble_new')
    assert np.array_equal(nc.variables['grid'][:], np.arange(6))
    nc.close()
    back = load_netcdf(path, array=True)
    assert back.dims == w.dims and np.array_equal(back.values, w.values)
    assert isinstance(back.indexes['grid'], pd.MultiIndex) and back.indexes['grid'].equals(mi)
    assert list(back.indexes['ensemble']) == [0, 1, 2, 3]
    # global weights, plain integer grid: nothing to encode
    w2 = xrlite.DataArray(rng.normal(size=(3, 3)), dict(ensemble=np.arange(3), ensemble_new=np.arange(3)), ('ensemble', 'ensemble_new'))
    save_netcdf(w2, path)
    assert np.array_equal(load_netcdf(path, array=True).values, w2.values)
    alg = ETKF(weight_save_path=path)
    st = xrlite.DataArray(np.zeros((1, 1, 3, 5)), dict(ensemble=np.arange(3), grid=np.arange(5)), ('var_name', 'time', 'ensemble', 'grid'))
    assert np.array_equal(alg._weights_through_store(st, w2.values), w2.values)          # filter.py:159-162: store, load back
    w3 = rng.normal(size=(5, 3, 3))
    assert np.array_equal(alg._weights_through_store(st, w3), w3)
    assert alg.load_weights().dims == ('grid', 'ensemble', 'ensemble_new')
    with pytest.raises(NotImplementedError):
        load_netcdf(path)


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


# ---- observation operators that select grid columns ---------------------------------------------------------
def _operator_objects(n_grid=40, k=6, n_time=3, seed=3, grid=None):
    rng = np.random.RandomState(seed)
    t = pd.to_datetime("1992-12-25") + pd.to_timedelta(np.arange(n_time), unit="h")
    grid = np.arange(n_grid) if grid is None else grid
    state = xrlite.DataArray(rng.normal(size=(2, n_time, k, n_grid)),
                             dict(var_name=["y", "x"], time=t, ensemble=np.arange(k), grid=grid),
                             ("var_name", "time", "ensemble", "grid"))
    return rng, t, state


def _obs_for(rng, times, obs_grid, var=0.25):
    n_o = len(obs_grid)
    return xrlite.Dataset({
        "observations": xrlite.DataArray(rng.normal(size=(len(times), n_o)), dict(time=times, obs_grid_1=obs_grid), ("time", "obs_grid_1")),
        "covariance": xrlite.DataArray(np.full(n_o, var) + 0.01 * np.arange(n_o), dict(obs_grid_1=obs_grid), ("obs_grid_1",))})


def test_identity_operator_semantics():
    """obs_ops/lorenz_96/identity.py:74-92 and obs_ops/base_ops.py:63-75."""
    from pytassim_b200.obs_ops import IdentityOperator
    rng, t, state = _operator_objects()
    op_all = IdentityOperator(len_grid=40)                                     # None -> every grid point
    obs = _obs_for(rng, t[[2, 0]], np.arange(40))                              # observation times in a different order
    hx = op_all(obs, state)
    assert hx.dims == ("time", "ensemble", "obs_grid_1") and hx.shape == (2, 6, 40)
    np.testing.assert_array_equal(hx.values, state.values[1][[2, 0]])         # var 'x' is position 1
    assert list(hx.indexes["time"]) == list(t[[2, 0]])
    op_list = IdentityOperator(obs_points=[7, 3, 39], len_grid=40)
    obs3 = _obs_for(rng, t[:1], [7, 3, 39])
    np.testing.assert_array_equal(op_list(obs3, state).values, state.values[1][:1][:, :, [7, 3, 39]])
    op_rand = IdentityOperator(obs_points=5, len_grid=40, random_state=np.random.RandomState(0))
    expect = np.random.RandomState(0).choice(40, size=5, replace=False)
    np.testing.assert_array_equal(op_rand.device_index(_obs_for(rng, t[:1], expect), state)[2], expect)
    # labels, not positions: a grid index with shuffled labels
    perm = np.random.RandomState(1).permutation(40)
    _, _, shuffled = _operator_objects(grid=perm)
    pos = op_list.device_index(obs3, shuffled)[2]
    np.testing.assert_array_equal(perm[pos], [7, 3, 39])
    with pytest.raises(KeyError):                                              # a grid label the state does not have
        IdentityOperator(obs_points=[41], len_grid=40).device_index(_obs_for(rng, t[:1], [41]), state)
    with pytest.raises(KeyError):                                              # an observation time the state does not have
        op_all(_obs_for(rng, t[:1] + pd.Timedelta("30min"), np.arange(40)), state)
    with pytest.raises(ValueError):                                            # obs_grid_1 of the wrong length
        op_all(_obs_for(rng, t[:1], np.arange(39)), state)


def test_nearest_grid_operator_semantics():
    """examples/benchmark_letkf.py:90-104: linspace positions, nearest label, ties to the larger label (pandas)."""
    from pytassim_b200.obs_ops import NearestGridOperator
    rng, t, state = _operator_objects(n_grid=40)
    op = NearestGridOperator(len_grid=40, nr_obs=7)
    np.testing.assert_allclose(op.obs_grid, np.linspace(0, 40, 7, endpoint=False))
    pos = op.device_index(_obs_for(rng, t[:1], op.obs_grid), state)[2]
    np.testing.assert_array_equal(pos, [0, 6, 11, 17, 23, 29, 34])
    ties = NearestGridOperator(len_grid=40, obs_grid=[0.5, 1.5, 2.49, 2.51, -3.0, 99.0])
    np.testing.assert_array_equal(ties.grid_positions(state.indexes["grid"]), [1, 2, 2, 3, 0, 39])
    op1k = NearestGridOperator(len_grid=10_000, nr_obs=1_000)                  # the benchmark script's default shape
    np.testing.assert_array_equal(op1k.grid_positions(pd.Index(np.arange(10_000))), np.arange(0, 10_000, 10))


def test_gather_offsets_equal_host_operator_stack():
    """The device gather plan (src offsets into the flattened pseudo state) reproduces operator -> _stack_obs: same HX,
    observations, variances and obs_info, dataset-major / time-major / obs_grid_1-minor (base.py:223-241)."""
    from pytassim_b200.obs_ops import IdentityOperator, NearestGridOperator
    rng, t, state = _operator_objects(n_grid=40, k=6, n_time=3)
    ds1 = _obs_for(rng, t[[0, 2]], [5, 9, 33]); ds1.obs.operator = IdentityOperator(obs_points=[5, 9, 33], len_grid=40)
    op2 = NearestGridOperator(len_grid=40, nr_obs=7)
    ds2 = _obs_for(rng, t[[1]], op2.obs_grid); ds2.obs.operator = op2
    obs = (ds1, ds2)
    src, stride, y, var, info = BaseAssimilation._stack_gather_inputs(state, obs)
    ens_obs, filtered = BaseAssimilation._apply_obs_operator(state, obs)
    hx_ref, y_ref, var_ref, info_ref = BaseAssimilation._stack_obs_space_inputs(ens_obs, filtered)
    flat = state.values.reshape(-1)
    hx = np.stack([flat[src + i * stride] for i in range(6)])
    np.testing.assert_array_equal(hx, hx_ref)
    np.testing.assert_array_equal(y, y_ref); np.testing.assert_array_equal(var, var_ref); np.testing.assert_array_equal(info, info_ref)
    assert stride == 40 and src.dtype == np.int64 and src.shape == (2 * 3 + 7,)
    # a dataset with a plain callable operator or a correlated R switches the whole call to the host operators
    ds3 = _obs_for(rng, t[[1]], np.arange(40)); ds3.obs.operator = lambda o, s: IdentityOperator(len_grid=40)(o, s)
    assert BaseAssimilation._stack_gather_inputs(state, (ds1, ds3)) is None


def test_forward_model_builds_the_pseudo_state(golden):
    """base.py:327-357 / filter.py:139-146: without an explicit pseudo state a given forward model is run on
    mean + perturbations (identity prior weights) and its second return value is the pseudo state."""
    _, state, _ = _fixture_objects(golden)
    seen = {}

    def model(model_state, iter_num):
        seen["state"], seen["iter"] = model_state, iter_num
        return model_state, model_state.copy(data=model_state.values * 2.0)
    alg = ETKF(forward_model=model)
    pseudo = alg.get_pseudo_state(None, state)
    mean = state.values.mean(axis=2, keepdims=True)
    np.testing.assert_array_equal(seen["state"].values, mean + (state.values - mean))
    np.testing.assert_array_equal(pseudo.values, seen["state"].values * 2.0)
    assert seen["iter"] == 0 and pseudo.dims == state.dims
    assert alg.get_pseudo_state(state, state) is state                     # an explicit pseudo state wins
    assert ETKF().get_pseudo_state(None, state) is state                   # no model: the state itself

    def bad_model(model_state, iter_num):
        return model_state, xrlite.DataArray(model_state.values, {}, ("v", "time", "ensemble", "grid"))
    with pytest.raises(StateError):
        ETKF(forward_model=bad_model).get_pseudo_state(None, state)


def test_position_operator():
    from pytassim_b200.obs_ops import PositionOperator
    rng, t, state = _operator_objects()
    op = PositionOperator([3, 3, 39])
    obs = _obs_for(rng, t[:1], [0., 1., 2.])
    np.testing.assert_array_equal(op(obs, state).values, state.values[1][:1][:, :, [3, 3, 39]])
    with pytest.raises(IndexError):
        PositionOperator([40])(_obs_for(rng, t[:1], [0.]), state)


def test_kernelised_classes_signatures_and_kernel_check():
    from pytassim_b200.interface import KETKF, LKETKF
    from pytassim_b200.kernels import LinearKernel
    alg = LKETKF(localization=None, kernel=LinearKernel(), inf_factor=1.0, smoother=False, gpu=False, pre_transform=None,
                 post_transform=None, chunksize=10, weight_save_path=None, forward_model=None)      # lketkf.py:79-91
    assert isinstance(alg.kernel, LinearKernel) and alg.chunks == {"grid": 10}
    glob = KETKF(kernel=LinearKernel(), inf_factor=1.5, smoother=True, gpu=False, pre_transform=None, post_transform=None,
                 weight_save_path=None, forward_model=None)                                          # ketkf.py:69-78
    assert str(glob) == "Global KETKF(inf_factor=1.5, kernel=LinearKernel)" and glob.smoother
    with pytest.raises(NotImplementedError):
        KETKF(kernel=lambda x, y: x @ y.T)


def test_kernel_descriptors_compile_to_programs():
    """pytassim_b200.kernels: reference constructor signatures and composition operators (kernels/base_kernels.py:40-161)
    compile to the postfix program of b200da_plan_set_kernel; the float32 quirks of the reference are mirrored."""
    import torch
    from pytassim_b200 import _cabi, kernels as K
    from pytassim_b200.interface import KETKF, LKETKF
    from pytassim_b200.testing import kernel_cases as kc
    assert K.GaussKernel(lengthscale=2.).program() == [(_cabi.KOP_GAUSS, 2.0, 0.0)]
    # rbf.py:103-104: the length scale is computed in the parameter's own type (float32 for the reference's tensor default)
    ls32 = K.RBFKernel(gamma=torch.tensor(0.3)).program()[0][1]
    assert ls32 == float((0.5 / torch.tensor(0.3)) ** 0.5) and ls32 != (0.5 / 0.3) ** 0.5
    assert K.RBFKernel(gamma=0.3).program()[0][1] == (0.5 / 0.3) ** 0.5
    # scale.py:70-72: the constant is float32-rounded
    assert K.ScaleKernel(0.01).program() == [(_cabi.KOP_SCALE, float(np.float32(0.01)), 0.0)]
    comp = K.GaussKernel(3.) * K.ScaleKernel(2.) + K.DiagKernel(0.5)
    assert [op for op, _, _ in comp.program()] == [_cabi.KOP_GAUSS, _cabi.KOP_SCALE, _cabi.KOP_MUL, _cabi.KOP_DIAG, _cabi.KOP_ADD]
    assert str(comp) == "GaussKernel(l=3.0)*ScaleKernel(2.0)+DiagKernel(0.5)" and comp.positive_semidefinite
    assert not K.TanhKernel().positive_semidefinite and not (K.GaussKernel() ** K.ScaleKernel(2.)).positive_semidefinite
    assert K.PolyKernel(2., 1.).positive_semidefinite and not K.PolyKernel(2.5, 1.).positive_semidefinite
    for name, build in kc.KERNEL_CASES:
        prog = build(K, 40).program()
        depth = 0
        for op, _, _ in prog:
            depth += -1 if op >= _cabi.KOP_ADD else 1
        assert depth == 1 and len(prog) <= _cabi.MAX_KERNEL_OPS, name
    for cls in (K.OrnsteinUhlenbeckKernel, K.PeriodicKernel):         # L1 kernels are not functions of the Gram
        with pytest.raises(NotImplementedError):
            cls(1.0)
    with pytest.raises(NotImplementedError):
        K.GaussKernel() + (lambda x, y: x @ y.T)
    alg = KETKF(kernel=K.RBFKernel(gamma=0.1), inf_factor=1.2)
    assert str(alg) == "Global KETKF(inf_factor=1.2, kernel=RBFKernel(γ=0.1))" and alg._kernel_key() == tuple(alg.kernel.program())
    assert LKETKF(kernel=K.LinearKernel())._kernel_key() == ()
    alg.kernel = K.PolyKernel()                                        # ketkf.py:118-123: the setter swaps the core module
    assert alg._kernel_key() == ((_cabi.KOP_POLY, 2.0, 1.0),) and alg._engines == {}


def test_ienks_classes_signatures_and_bounds():
    """interface/ienks.py:34-164, interface/lienks.py:40-163: constructor signatures, string forms, tau / epsilon bounds
    (utilities/decorators.py:51-75) and the bundle's model weights (ienks.py:153-160)."""
    import torch
    from pytassim_b200.interface import IEnKSTransform, IEnKSBundle, LocalizedIEnKSTransform, LocalizedIEnKSBundle
    fm = lambda st, it: (st, st)
    alg = IEnKSTransform(forward_model=fm, tau=0.5, max_iter=4, smoother=False, gpu=False, pre_transform=None,
                         post_transform=None, weight_save_path=None)
    assert str(alg) == "IEnKSTransform(tau=0.5)" and repr(alg) == "IEnKSTransform(0.5)" and alg.max_iter == 4
    bun = IEnKSBundle(forward_model=fm, tau=1.0, epsilon=1E-4, max_iter=10)
    assert str(bun) == "IEnKSBundle(epsilon=0.0001, tau=1.0)" and repr(bun) == "IEnKSBundle(0.0001,1.0)"
    loc = GaspariCohn((10.,), AbsDistance1D())
    lt = LocalizedIEnKSTransform(forward_model=fm, localization=loc, tau=1.0, max_iter=10, smoother=False, gpu=False,
                                 pre_transform=None, post_transform=None, chunksize=10, weight_save_path=None)
    assert str(lt) == "Localized IEnKSTransform(loc=GaspariCohn(l=[10.]), tau=1.0)" and lt.chunks == {"grid": 10}
    lb = LocalizedIEnKSBundle(forward_model=fm, localization=loc, tau=0.9, epsilon=1E-2)
    assert repr(lb) == "LIEnKSBundle(GaspariCohn,0.01,0.9)"
    for bad in (-0.1, 1.5):
        with pytest.raises(ValueError):
            alg.tau = bad
    with pytest.raises(ValueError):
        bun.epsilon = -1.0
    w = torch.arange(9, dtype=torch.float64).reshape(3, 3)
    bun.epsilon = 0.5
    np.testing.assert_array_equal(bun._get_model_weights(w).numpy(), 0.5 * np.eye(3) + w.numpy().mean(axis=1, keepdims=True))
    assert alg._get_model_weights(w) is w


def test_product_distance_equals_oracle_rows():
    from pytassim_b200.localization import ProductDistance, HaversineDistance
    rng = np.random.RandomState(3)
    obs = np.column_stack([np.zeros(50), rng.uniform(-80, 80, 50), rng.uniform(0, 360, 50), rng.uniform(0, 10, 50), rng.uniform(0, 5, 50)])
    grid = np.array([0.0, 12.0, 200.0, 3.0, 1.0])
    m = ProductDistance(HaversineDistance(6371.0), n_extra=2)
    ref = orc.make_dist_product(orc.make_dist_haversine(6371.0), 2, 2)(grid, obs)
    np.testing.assert_array_equal(m(grid, pd.DataFrame(obs)), ref)
    assert m.n_coord == 4 and m.n_extra == 2 and m.metric_id == HaversineDistance().metric_id and ref.shape == (3, 50)
    m1 = ProductDistance(PeriodicDistance1D(12.0))
    np.testing.assert_array_equal(m1(grid[:3], obs[:, :3]), orc.make_dist_product(orc.make_dist_periodic1d(12.0), 1, 1)(grid[:3], obs[:, :3]))
    with pytest.raises(ValueError):
        ProductDistance(m1)
    with pytest.raises(ValueError):
        ProductDistance(AbsDistance1D(), n_extra=3)


def test_netcdf_store_coordinate_encodings(tmp_path):
    """netCDF-3 has no 64-bit integers, booleans or datetimes: the store applies the casts xarray's netCDF-3 encoder applies
    (int64 -> int32, datetime64 -> seconds since the epoch with units) and undoes the time encoding on load; a named data
    variable (a DataArray saved with a name by the reference) is found as the only non-coordinate variable."""
    from scipy.io import netcdf_file
    from pytassim_b200.utilities import save_netcdf, load_netcdf
    t = pd.to_datetime("1992-12-25") + pd.to_timedelta(np.arange(3), unit="h")
    arr = xrlite.DataArray(np.arange(12.0).reshape(3, 4), dict(time=t, grid=np.arange(4, dtype=np.int64) * 10), ("time", "grid"))
    path = str(tmp_path / "a.nc")
    save_netcdf(arr, path)
    nc = netcdf_file(path, "r", mmap=False)
    assert nc.variables["grid"][:].dtype.itemsize == 4 and nc.variables["time"].units.decode().startswith("seconds since 1970")
    nc.close()
    back = load_netcdf(path, array=True)
    assert list(pd.DatetimeIndex(back.indexes["time"])) == list(t) and list(back.indexes["grid"]) == [0, 10, 20, 30]
    assert np.array_equal(back.values, arr.values)
    with pytest.raises(ValueError):                                  # does not fit 32-bit integers
        save_netcdf(xrlite.DataArray(np.zeros(2), dict(grid=np.array([0, 2 ** 40])), ("grid", )), path)
    # string labels: character arrays with a trailing string<n> dimension, as xarray's netCDF-3 encoder writes them
    save_netcdf(xrlite.DataArray(np.arange(6.0).reshape(2, 3), dict(var_name=["x", "temperature"], grid=np.arange(3)),
                                 ("var_name", "grid")), path)
    lab = load_netcdf(path, array=True)
    assert list(lab.indexes["var_name"]) == ["x", "temperature"] and np.array_equal(lab.values, np.arange(6.0).reshape(2, 3))
    nc = netcdf_file(path, "w", version=2)                           # a file as xarray writes a NAMED DataArray
    nc.createDimension("ensemble", 2); nc.createDimension("ensemble_new", 2)
    v = nc.createVariable("ensemble", "i4", ("ensemble", )); v[:] = [0, 1]
    v = nc.createVariable("weights", "f8", ("ensemble", "ensemble_new")); v[:] = np.eye(2)
    nc.close()
    named = load_netcdf(path, array=True)
    assert named.dims == ("ensemble", "ensemble_new") and np.array_equal(named.values, np.eye(2))


def test_netcdf_store_beyond_the_fixed_variable_limit(tmp_path, monkeypatch):
    """Per-grid-point weights of 2 GiB or more (N ~ 168 000 at k = 40) exceed scipy's signed 32-bit ``vsize`` of a fixed
    netCDF-3 variable: they are written with ``grid`` as the record dimension.  The limit is lowered here so that the same
    code path runs on a small array; the file must round-trip, MultiIndex included, and a 1-D array over the limit (no
    dimension to make records of) must raise a clear error instead of struct.error."""
    from scipy.io import netcdf_file
    from pytassim_b200.utilities import netcdf as store
    rng = np.random.RandomState(1)
    mi = pd.MultiIndex.from_product((np.arange(7) * 0.5, [0.0, 180.0]), names=['lat', 'lon'])
    w = xrlite.DataArray(rng.normal(size=(14, 5, 5)), dict(grid=mi, ensemble=np.arange(5), ensemble_new=np.arange(5)),
                         ('grid', 'ensemble', 'ensemble_new'))
    monkeypatch.setattr(store, "FIXED_VARIABLE_LIMIT", 1000)          # 14 x 5 x 5 x 8 = 2800 bytes > limit
    path = str(tmp_path / "big.nc")
    store.save_netcdf(w, path)
    nc = netcdf_file(path, 'r', mmap=False)
    assert nc.dimensions['grid'] is None and nc.variables[store.DATAARRAY_VARIABLE].isrec      # record layout
    nc.close()
    back = store.load_netcdf(path, array=True)
    assert np.array_equal(back.values, w.values) and back.indexes['grid'].equals(mi)
    with pytest.raises(ValueError, match="2 GiB"):
        store.save_netcdf(xrlite.DataArray(np.zeros(400), dict(grid=np.arange(400)), ('grid', )), path)
    # the real limit: sizes are checked up front, nothing of this size is allocated here
    assert store.FIXED_VARIABLE_LIMIT == 1000 and 168_000 * 40 * 40 * 8 > 2 ** 31 - 4


def _emulate_program(prog, xy, xx, yy, diag, same_set):
    """numpy emulation of the postfix interpreter of csrc/kernelise_kernel.cuh (kernel_eval) — the semantics of the program the
    descriptors compile to, checked here without a GPU against the oracle's restatement of pytassim/kernels."""
    from pytassim_b200 import _cabi as c
    st = []
    for op, a, b in prog:
        if op == c.KOP_LINEAR:
            st.append(xy)
        elif op == c.KOP_GAUSS:
            st.append(np.exp(-(np.maximum((xx + yy - 2.0 * xy) / (a * a), 0.0) / 2.0)))
        elif op == c.KOP_POLY:
            st.append(np.power(xy + b, a))
        elif op == c.KOP_TANH:
            st.append(np.tanh(a * xy + b))
        elif op == c.KOP_RATIONAL:
            st.append(np.power(1.0 + np.maximum((xx + yy - 2.0 * xy) / (a * a), 0.0) / (2.0 * b), -b))
        elif op == c.KOP_SCALE:
            st.append(np.full_like(xy, a))
        elif op == c.KOP_DIAG:
            st.append(np.where(diag & same_set, a, 0.0))
        else:
            y, x = st.pop(), st.pop()
            st.append(x + y if op == c.KOP_ADD else x * y if op == c.KOP_MUL else np.power(x, y))
    assert len(st) == 1
    return st[0]


def test_kernel_programs_reproduce_reference_kernels_from_the_gram(golden):
    """Every kernel configuration: descriptor -> program -> (emulated) evaluation on entries of the augmented Gram, double
    centring as k_kernelise does it, weights by the oracle's evd route == the reference's ``KETKFModule(kernel)`` output
    (tests/golden/ketkf_kernels.npz).  This is the algebra the device path relies on: kernels as functions of the Gram."""
    from pytassim_b200 import kernels as K
    from pytassim_b200.testing import kernel_cases as kc
    g = golden("ketkf_kernels.npz")
    for i, (k, p, rho) in enumerate(kc.PROBLEM_SIZES):
        perts, obs = g["c%d_perts" % i], g["c%d_obs" % i]
        aug = np.concatenate([perts, obs], axis=0)
        gram = aug @ aug.T
        dg = np.diag(gram)
        for name, build in kc.KERNEL_CASES:
            prog = build(K, p).program()
            kp = _emulate_program(prog, gram[:k, :k], dg[:k, None], dg[None, :k], np.eye(k, dtype=bool), True)
            ko = _emulate_program(prog, gram[:k, k:], dg[:k, None], dg[None, k:], np.zeros((k, 1), dtype=bool), False)
            rmean = kp.mean(axis=1, keepdims=True)
            mu = rmean.mean()
            kc_mat = kp - rmean.T - (rmean - mu)                                     # core/ketkf.py:81-85
            ko_c = ko - ko.mean() - (rmean - mu)                                    # :91-92
            evals, evects, evals_inv = orc.evd(kc_mat, (k - 1) / rho)
            w = orc.rev_evd(evals_inv, evects) @ ko_c + orc.rev_evd(np.sqrt((k - 1) * evals_inv), evects)
            np.testing.assert_allclose(w, g["c%d_w_%s" % (i, name)], rtol=1e-10, atol=1e-10, err_msg="%s k=%d" % (name, k))


def test_ienks_gram_reformulation_reproduces_reference(golden):
    """The algebra behind k_ienks_pre (csrc/ienks_kernel.cuh), in numpy: from the Gram C = Yn Yn^T, b = Yn d^T and the incoming
    weights, A' = (1 - tau)(k - 1) T^T T + tau S C S^T, b' = P w - tau grad, then the ETKF solve with inflation 1 / tau — equals
    the reference's SVD-based ``IEnKSTransformModule`` / ``IEnKSBundleModule`` over three iterations (tests/golden/ienks.npz)."""
    g = golden("ienks.npz")

    def step(w_in, yn, d, tau, eps):
        k = w_in.shape[0]
        c, b = yn @ yn.T, yn @ d.ravel()
        w = (w_in - np.eye(k)).mean(axis=1)                                         # core/ienks.py:50-51
        t = np.linalg.inv(w_in - w[:, None])                                        # :62-65
        scs, sb = (t @ c @ t.T, t @ b) if eps is None else (c / eps / eps, b / eps)   # :74-75 / :173
        a = (1 - tau) * (k - 1) * t.T @ t + tau * scs
        a = 0.5 * (a + a.T)
        alpha = tau * (k - 1)
        b_slot = (a + alpha * np.eye(k)) @ w - tau * ((k - 1) * w - sb)
        evals, evects, evals_inv = orc.evd(a, alpha)                                # the solve kernels: rho = 1 / tau
        return (orc.rev_evd(evals_inv, evects) @ b_slot)[:, None] + orc.rev_evd(np.sqrt((k - 1) * evals_inv), evects)
    for i in range(int(g["n_cases"])):
        perts, obs, tau = g["c%d_perts" % i], g["c%d_obs" % i], float(g["c%d_tau" % i])
        for variant, eps in (("transform", None), ("bundle", float(g["c%d_eps" % i]))):
            w = np.eye(perts.shape[0])
            for it in range(3):
                w = step(w, perts * (1.0 if eps is None else eps), obs, tau, eps)
                np.testing.assert_allclose(w, g["c%d_%s_w%d" % (i, variant, it)], rtol=1e-10, atol=1e-10)


def test_inplace_gauss_jordan_as_in_ienks_pre():
    """numpy emulation of the inversion inside k_ienks_pre (csrc/ienks_kernel.cuh): in-place Gauss-Jordan with row pivoting,
    the pivot chosen on the upper 32 bits of |a| (any pivot within 2^-20 of the largest), the unit column folded into the
    scaled pivot row, the moved row read from its staging copy, and the row swaps undone as column swaps in reverse order."""
    def invert(a):
        a = a.copy()
        k = len(a)
        pivs = np.zeros(k, dtype=int)
        for c in range(k):
            keys = (np.abs(a[c:, c]).view(np.uint64) >> np.uint64(32)).astype(np.uint32)   # __double2hiint(fabs(.))
            pr = c + int(np.argmax(keys))
            pivs[c] = pr
            pinv = 1.0 / a[pr, c]
            rowa = np.where(np.arange(k) == c, 1.0, a[pr]) * pinv
            rowc, colv = a[c].copy(), a[:, c].copy()
            for r in range(k):
                if r == c:
                    a[r] = rowa
                    continue
                moved = r == pr
                src = rowc if moved else a[r]
                f = rowc[c] if moved else colv[r]
                a[r] = np.where(np.arange(k) == c, 0.0, src) - f * rowa
        for c in range(k - 1, -1, -1):
            p = pivs[c]
            if p != c:
                a[:, [c, p]] = a[:, [p, c]]
        return a
    rng = np.random.RandomState(8)
    for k in (2, 3, 10, 40, 50, 96):
        m_ = rng.normal(size=(k, k))
        np.testing.assert_allclose(invert(m_) @ m_, np.eye(k), atol=1e-9 * np.linalg.cond(m_))
        np.testing.assert_allclose(invert(m_), np.linalg.inv(m_), rtol=0, atol=1e-12 * np.linalg.cond(m_) * np.abs(np.linalg.inv(m_)).max())
    perm = np.eye(5)[[3, 0, 4, 1, 2]] * np.array([2., -3., 0.5, 4., 1.])              # needs a swap at every step
    np.testing.assert_allclose(invert(perm), np.linalg.inv(perm), atol=1e-15)


def test_newton_schulz_solve_as_in_the_solve_kernel():
    """numpy emulation of csrc/ns_solve_kernel.cuh against the reference's eigendecomposition route (core/etkf.py:57-77,102;
    interface/base.py:257-278): scaled coupled Newton-Schulz iteration with the tracked spectral bound, every product kept as its
    lower triangle (symmetric storage), sqrt(g) folded into the stored T, the dense result left unscaled, and the update applied
    as  mean + xp . w_mean + sqrt(k - 1) zs (Z xp)  without forming W."""
    def sym(m_):
        return np.tril(m_) + np.tril(m_, -1).T

    def inv_sqrt(a_mat, alpha, conv=2e-8):
        k = len(a_mat)
        s = min(np.linalg.norm(a_mat), np.abs(a_mat).sum(axis=1).max()) * (1.0 + 1e-12)
        lo = min(alpha / s, 1.0)
        g = 3.0 / (1.0 + np.sqrt(lo) + lo)
        sg = np.sqrt(g)

        def next_scaling(g, lo):
            m_ = g * lo
            lo = min(1.0, 0.25 * m_ * (3.0 - m_) ** 2)
            g = 3.0 / (1.0 + np.sqrt(lo) + lo)
            return g, np.sqrt(g), lo
        y = a_mat / s
        z = sg * (1.5 * np.eye(k) - 0.5 * g * y)                  # iteration 0: Z1 = sqrt(g) T, Y1 = Y0 Z1
        g, sg, lo = next_scaling(g, lo)
        y = sym(y @ z)
        iters = 1
        while iters < 64:
            last = (1.0 - lo) < conv
            gc, sgc = g, sg
            g, sg, lo = next_scaling(g, lo)
            t = sym(-0.5 * gc * sgc * (z @ y)) + 1.5 * sgc * np.eye(k)
            z = sym(t @ z)
            iters += 1
            if last:
                break
            y = sym(y @ t)
        return z, np.sqrt(1.0 / s), iters

    rng = np.random.RandomState(11)
    for k, p, scale, rho in ((40, 60, 1.0, 1.1), (50, 500, 0.3, 1.05), (16, 8, 2.0, 1.0), (24, 300, 1.5, 1.2)):
        perts = rng.normal(size=(k, p)) * scale
        perts -= perts.mean(axis=0)
        obs = rng.normal(size=(1, p)) * scale
        alpha = (k - 1) / rho
        z, zs, iters = inv_sqrt(perts @ perts.T + alpha * np.eye(k), alpha)
        assert iters <= 10
        b = perts @ obs[0]
        w_mean = zs * zs * (z @ (z @ b))
        w_ref = orc.etkf_weights(perts, obs, rho)
        np.testing.assert_allclose(w_mean[:, None] + np.sqrt(k - 1) * zs * z, w_ref, rtol=0, atol=1e-11 * np.abs(w_ref).max())
        x = rng.normal(size=(1, 1, k, 3)) + 280.0
        ref = orc.apply_weights(x, w_ref)
        for gp in range(3):
            col = x[0, 0, :, gp]
            mean = col.sum() / k
            xp = col - mean
            xa = mean + (xp @ w_mean + np.sqrt(k - 1) * zs * (z @ xp))
            np.testing.assert_allclose(xa, ref[0, 0, :, gp], rtol=1e-12, atol=0)


class _OracleIenksEngine(object):
    """CPU stand-in for the engine behind ``VarAssimilation.update_state`` (test infrastructure: the oracle does the numerics)
    so that the control flow of the outer loop (interface/variational.py:105-135) is exercised without a GPU."""
    device = torch.device("cpu")

    def apply_weights(self, x, weights):
        return torch.as_tensor(orc.apply_weights(x.numpy()[:, None], weights.numpy())[:, 0])

    def ienks_weights(self, weights, perts, innov, tau=1.0, epsilon=None):
        return torch.as_tensor(orc.ienks_weights(weights.numpy(), np.asarray(perts), np.asarray(innov)[None], tau, epsilon))


def test_ienks_outer_loop_control_flow(golden, monkeypatch):
    """tests/unit_tests/interface/test_ienks.py:143-161, 215-237 of the reference: a given pseudo state skips the first model
    propagation, later iterations propagate; max_iter = 1 with an identity model equals the ETKF without inflation; smoother
    mode propagates the analysis once more (variational.py:132-133); the bundle variant propagates epsilon * I + mean weights."""
    from pytassim_b200.interface import IEnKSTransform, IEnKSBundle
    g, state, obs = _fixture_objects(golden)
    ob_last = obs.isel(time=[2])
    t_last = "1992-12-25 02:00"
    monkeypatch.setattr(IEnKSTransform, "_analysis_engine", lambda self, st, k, n_slices: _OracleIenksEngine())
    calls = []

    def model(st, iter_num):
        calls.append((iter_num, np.asarray(st.values).copy()))
        return st, st
    st_last = state.isel(time=[2])
    alg = IEnKSTransform(model, tau=1.0, max_iter=1)
    ana = alg.assimilate(state, ob_last, st_last, analysis_time=t_last)            # pseudo state given: no propagation
    assert calls == []
    innov, perts = orc.obs_space_variables([g["state"][0, 2][:, None, :]], [g["obs"][2:3]], [g["cov"]])
    ref = orc.apply_weights(g["state"][:, 2:3], orc.etkf_weights(perts, innov, 1.0))
    np.testing.assert_allclose(ana.values, ref, rtol=1e-10, atol=1e-10)           # == ETKF (test_ienks.py:215-237)
    assert list(ana.indexes["time"]) == [pd.Timestamp(t_last)]
    alg.max_iter = 2
    alg.assimilate(state, ob_last, st_last, analysis_time=t_last)
    assert [c[0] for c in calls] == [1]                                             # test_ienks.py:153-161
    del calls[:]
    alg.smoother = True
    alg.assimilate(state, ob_last, None, analysis_time=t_last)
    assert [c[0] for c in calls] == [0, 1, 2]                                       # two iterations + the smoother propagation
    np.testing.assert_allclose(calls[0][1], g["state"][:, 2:3], rtol=0, atol=1e-13)  # prior weights = identity
    del calls[:]
    bun = IEnKSBundle(model, tau=1.0, epsilon=1e-2, max_iter=1)
    bun.assimilate(state, ob_last, None, analysis_time=t_last)
    mean = g["state"][:, 2:3].mean(axis=2, keepdims=True)
    np.testing.assert_allclose(calls[0][1], mean + 1e-2 * (g["state"][:, 2:3] - mean), rtol=0, atol=1e-13)   # ienks.py:153-160
    with pytest.raises(KeyError):
        alg.update_state(state, [ob_last], None, pd.Timestamp("2000-01-01"))


def test_kernelised_interface_configures_engines(golden, monkeypatch):
    """KETKF / LKETKF hand their kernel to every engine they create and drop cached engines when the kernel changes
    (interface/ketkf.py:118-123 swaps the core module); the plain ETKF / LETKF never set a program."""
    from pytassim_b200 import kernels as K
    from pytassim_b200.interface import KETKF, etkf as etkf_mod
    created = []

    class StubEngine(object):
        def __init__(self, k, n_slices, metric, radius, **kwargs):
            self.k, self.kernel_set = k, []
            created.append(self)

        def set_kernel(self, kernel):
            self.kernel_set.append(kernel)
            return self
    monkeypatch.setattr(etkf_mod, "LETKFEngine", StubEngine)
    alg = KETKF(kernel=K.RBFKernel(gamma=0.1), inf_factor=1.1)
    eng = alg._global_engine(10, 2)
    assert alg._global_engine(10, 2) is eng and len(created) == 1 and eng.kernel_set == [alg.kernel]
    alg.kernel = K.PolyKernel()
    eng2 = alg._global_engine(10, 2)
    assert eng2 is not eng and eng2.kernel_set == [alg.kernel]
    alg.inf_factor = 1.3                                         # etkf.py:93-97: a new core module -> new engines
    assert alg._global_engine(10, 2) is not eng2
    plain = ETKF(inf_factor=1.1)
    assert plain._global_engine(10, 2).kernel_set == [] and plain._kernel_key() == ()
    lin = KETKF()                                                # default LinearKernel: handed over, compiles to no program
    assert lin._global_engine(10, 2).kernel_set[0].is_linear and lin._kernel_key() == ()


def test_per_point_call_forms_with_stub_module():
    """interface/per_point.py against the reference's closures (interface/wrapper.py:54-62, 86-98) on stand-ins: the bridge
    returns the first argument's dtype, the localized form selects and scales every argument except the skipped ones."""
    from pytassim_b200.interface.per_point import NumpyBridge, LocalObservations
    seen = {}

    def core(*tensors):
        seen["args"] = tensors
        return tensors[0].sum() * torch.eye(2, dtype=tensors[0].dtype)
    bridge = NumpyBridge(core, torch.device("cpu"), torch.float64)
    out = bridge(np.ones((2, 3), dtype=np.float32), np.zeros((1, 3), dtype=np.float32))
    assert out.dtype == np.float32 and np.array_equal(out, 6.0 * np.eye(2))
    assert all(t.dtype == torch.float64 for t in seen["args"])

    class Loc(object):
        def localize_obs(self, grid_info, obs_info):
            assert grid_info == "row" and obs_info == "info"
            return np.array([True, False, True, True]), np.array([0.25, 0.0, 1.0, 0.04])
    calls = []
    lm = LocalObservations(lambda *a: calls.append(a) or "w", Loc())
    weights, perts, innov = np.eye(2), np.arange(8.0).reshape(2, 4), np.arange(4.0)[None]
    assert lm("row", weights, perts, innov, obs_info="info", args_to_skip=(0, )) == "w"
    w_arg, p_arg, i_arg = calls[0]
    assert w_arg is weights
    np.testing.assert_allclose(p_arg, perts[:, [0, 2, 3]] * np.array([0.5, 1.0, 0.2]))
    np.testing.assert_allclose(i_arg, innov[:, [0, 2, 3]] * np.array([0.5, 1.0, 0.2]))
    assert LocalObservations(lambda *a: a, None)("row", 1, 2, obs_info=None) == (1, 2)        # wrapper.py:87: no localization


# ---- ambiguity protocol: the host decision rule is the reference's expression ----------------------------------------
def test_host_decision_is_bit_equal_to_the_reference_expression(golden):
    """``BaseLocalization.host_decision`` (the rule applied to pairs the device cannot decide) against the oracle's
    restatement of gaspari_cohn.py:97-136 / :216-254, which tests/test_oracle.py pins to reference-generated goldens."""
    rnd = np.random.RandomState(0)
    obs_rows = np.stack([np.zeros(500), rnd.uniform(-30.0, 70.0, size=500)], axis=1)
    grid_row = np.array([0.0, 20.0])
    loc = GaspariCohn((10.0,), AbsDistance1D())
    use, w = loc.host_decision(grid_row, obs_rows)
    use_ref, w_ref = orc.gaspari_cohn_localize(orc.dist_abs1d(grid_row, obs_rows), (10.0,), 1e-5)
    np.testing.assert_array_equal(use, use_ref)
    np.testing.assert_array_equal(w, w_ref)
    loc = GaspariCohnInf(10.0, AbsDistance1D())
    use, w = loc.host_decision(grid_row, obs_rows)
    use_ref, w_ref = orc.gaspari_cohn_inf_localize(orc.dist_abs1d(grid_row, obs_rows), 10.0, 1e-5)
    np.testing.assert_array_equal(use, use_ref)
    np.testing.assert_array_equal(w, w_ref)
    # one pair at a time (how the protocol calls it) gives the same bits as the vector call
    for j in (3, 77, 256):
        u1, w1 = GaspariCohn((10.0,), AbsDistance1D()).host_decision(grid_row, obs_rows[j:j + 1])
        ua, wa = GaspariCohn((10.0,), AbsDistance1D()).host_decision(grid_row, obs_rows)
        assert u1[0] == ua[j] and w1[0] == wa[j]
