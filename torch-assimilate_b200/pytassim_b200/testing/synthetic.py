"""Seeded synthetic inputs for the BASELINE.json configurations (array level, no xarray).

Mirrors the shapes of the reference's ``examples/benchmark_letkf.py:107-149`` (state N(0,1), obs N(0,1),
R = I, identity / nearest-gridpoint observation operator) as specified in SURVEY.md section 8(d).
Every generator returns a dict with

    state        (n_var, n_time, k, N)   float64, grid fastest  (pytassim/state.py:114)
    normed_perts (k, M)                  obs-space ensemble perturbations times R^{-1/2}
    normed_obs   (M,)                    innovations times R^{-1/2}      (interface/base.py:359-379)
    grid_rows    (N, 1+nc)               [t_unix, coords...]             (interface/mixin_local.py:50-69)
    obs_rows     (M, 1+nc)               [time, coords...]               (interface/mixin_local.py:45-47)

The obs-space variables are formed exactly as the reference does: H(x) - mean_k H(x) and y - mean_k H(x),
both divided by sqrt(var) with var = 1.
"""
import numpy as np

__all__ = ["lorenz96_1d", "sphere_latlon", "global_etkf", "obs_space_from_hx"]


def obs_space_from_hx(hx, y, var=None):
    """hx (k, M) ensemble in observation space, y (M,) observations -> (normed_perts, normed_obs).
    Restates interface/base.py:367-372 for one uncorrelated obs dataset (observation.py:241-245)."""
    mean = hx.mean(axis=0, keepdims=True)
    perts = hx - mean
    innov = y - mean[0]
    if var is not None:
        rcinv = 1 / np.sqrt(var)
        perts = perts * rcinv
        innov = innov * rcinv
    return perts, innov


def lorenz96_1d(n_grid=40, k=50, obs_stride=1, seed=42, dtype=np.float64):
    """cfg1 (n_grid=40, k=50, stride 1) / cfg2 (n_grid=100000, k=40, stride 2): a ring of n_grid points,
    every ``obs_stride``-th observed, identity H, R = I."""
    rnd = np.random.RandomState(seed)
    state = rnd.normal(size=(1, 1, k, n_grid))
    obs_pos = np.arange(0, n_grid, obs_stride, dtype=np.float64)
    y = rnd.normal(size=(obs_pos.size,))
    hx = state[0, 0][:, ::obs_stride]
    perts, innov = obs_space_from_hx(hx, y)
    grid_rows = np.stack([np.zeros(n_grid), np.arange(n_grid, dtype=np.float64)], axis=1)
    obs_rows = np.stack([np.zeros(obs_pos.size), obs_pos], axis=1)
    return dict(state=state.astype(dtype), normed_perts=np.ascontiguousarray(perts).astype(dtype),
                normed_obs=innov.astype(dtype), grid_rows=grid_rows, obs_rows=obs_rows, period=float(n_grid))


def sphere_latlon(nlat=1000, nlon=1000, k=50, n_obs=2_500_000, seed=42, dtype=np.float64, n_slices=1):
    """cfg3 / cfg5: regular lat-lon grid (lat_i = -90 + (180/nlat)(i + 1/2), lon_j = (360/nlon) j, grid index
    = i * nlon + j), ``n_obs`` observations uniform on the sphere, H = nearest grid point, R = I."""
    rnd = np.random.RandomState(seed)
    n_grid = nlat * nlon
    dlat, dlon = 180.0 / nlat, 360.0 / nlon
    lat = -90.0 + dlat * (np.arange(nlat) + 0.5)
    lon = dlon * np.arange(nlon)
    glat = np.repeat(lat, nlon)
    glon = np.tile(lon, nlat)
    state = rnd.normal(size=(n_slices, 1, k, n_grid)).astype(dtype)
    olat = np.degrees(np.arcsin(rnd.uniform(-1.0, 1.0, size=n_obs)))
    olon = np.degrees(rnd.uniform(0.0, 2.0 * np.pi, size=n_obs))
    ilat = np.clip(np.floor((olat + 90.0) / dlat).astype(np.int64), 0, nlat - 1)
    ilon = np.round(olon / dlon).astype(np.int64) % nlon
    h_index = ilat * nlon + ilon
    y = rnd.normal(size=(n_obs,))
    hx = state[0, 0][:, h_index].astype(np.float64)
    perts, innov = obs_space_from_hx(hx, y)
    grid_rows = np.stack([np.zeros(n_grid), glat, glon], axis=1)
    obs_rows = np.stack([np.zeros(n_obs), olat, olon], axis=1)
    return dict(state=state, normed_perts=np.ascontiguousarray(perts).astype(dtype), normed_obs=innov.astype(dtype),
                grid_rows=grid_rows, obs_rows=obs_rows, h_index=h_index)


def global_etkf(n_state=10_000_000, k=100, obs_stride=10, seed=42, dtype=np.float64):
    """cfg4: one dense ensemble-space solve, every ``obs_stride``-th state element observed."""
    rnd = np.random.RandomState(seed)
    state = rnd.normal(size=(1, 1, k, n_state)).astype(dtype)
    hx = state[0, 0][:, ::obs_stride].astype(np.float64)
    y = rnd.normal(size=(hx.shape[1],))
    perts, innov = obs_space_from_hx(hx, y)
    return dict(state=state, normed_perts=np.ascontiguousarray(perts).astype(dtype), normed_obs=innov.astype(dtype))
