"""A very small stand-in for the parts of xarray that the LETKF / ETKF interface touches.

xarray (and dask, netCDF4) are not installed in the build image, so the interface classes duck-type their inputs:
anything with ``.values``, ``.dims`` and ``.indexes`` is a state, anything with ``['observations']``,
``['covariance']`` and ``.obs.operator`` is an observation set.  Real ``xarray.DataArray`` / ``xarray.Dataset``
objects satisfy that protocol (with the reference's ``.obs`` accessor registered, pytassim/observation.py:51); these
classes satisfy it too and are what the tests in this repository use.
"""
import numpy as np
import pandas as pd

__all__ = ["DataArray", "Dataset"]


def _as_index(values, name=None):
    if isinstance(values, pd.Index):
        return values
    return pd.Index(np.asarray(values), name=name)


class DataArray(object):
    def __init__(self, values, coords, dims):
        self.values = np.asarray(values)
        self.dims = tuple(dims)
        if self.values.ndim != len(self.dims):
            raise ValueError("values have {0} dims but {1} names were given".format(self.values.ndim, len(self.dims)))
        self.indexes = {d: _as_index(coords[d], d) for d in self.dims if d in coords}
        for d, n in zip(self.dims, self.values.shape):
            if d not in self.indexes:
                self.indexes[d] = pd.RangeIndex(n, name=d)
            if len(self.indexes[d]) != n:
                raise ValueError("coordinate {0} has the wrong length".format(d))

    @property
    def shape(self):
        return self.values.shape

    def __getitem__(self, dim):
        return self.indexes[dim]

    def isel(self, **indexers):
        values, coords = self.values, dict(self.indexes)
        for dim, idx in indexers.items():
            ax = self.dims.index(dim)
            idx = np.atleast_1d(idx)
            if idx.dtype.kind in "iu" and idx.size > 0 and idx[0] >= 0 and np.array_equal(idx, np.arange(idx[0], idx[0] + idx.size)):
                sl = [slice(None)] * values.ndim            # a contiguous run of positions: a view, no copy of the data
                sl[ax] = slice(int(idx[0]), int(idx[0]) + idx.size)
                values = values[tuple(sl)]
            else:
                values = np.take(values, idx, axis=ax)
            coords[dim] = self.indexes[dim][idx]
        return DataArray(values, coords, self.dims)

    def transpose(self, *dims):
        order = [self.dims.index(d) for d in dims]
        return DataArray(np.transpose(self.values, order), self.indexes, dims)

    def copy(self, data=None):
        return DataArray(self.values.copy() if data is None else data, self.indexes, self.dims)


class _ObsAccessor(object):
    """Mirror of the ``xarray.Dataset.obs`` accessor (pytassim/observation.py:51-299) — validity + operator."""

    def __init__(self, ds):
        self.ds = ds
        self.operator = self._no_operator

    @staticmethod
    def _no_operator(obs_ds, state):
        raise NotImplementedError('No observation operator is set!')        # observation.py:297-299

    @property
    def correlated(self):
        return 'obs_grid_2' in self.ds['covariance'].dims                    # observation.py:104-112

    @property
    def valid(self):                                                         # observation.py:114-239
        try:
            obs, cov = self.ds['observations'], self.ds['covariance']
        except KeyError:
            return False
        if obs.dims[-2:] != ('time', 'obs_grid_1'):
            return False
        n_t, n_o = obs.shape[-2:]
        ok = {('obs_grid_1',): (n_o,), ('time', 'obs_grid_1'): (n_t, n_o), ('obs_grid_1', 'obs_grid_2'): (n_o, n_o),
              ('time', 'obs_grid_1', 'obs_grid_2'): (n_t, n_o, n_o)}
        return cov.dims in ok and tuple(cov.shape) == ok[cov.dims]


class Dataset(object):
    def __init__(self, data_vars):
        self.data_vars = dict(data_vars)
        self.obs = _ObsAccessor(self)

    def __getitem__(self, name):
        return self.data_vars[name]

    def isel(self, **indexers):
        out = Dataset({k: (v.isel(**{d: i for d, i in indexers.items() if d in v.dims})) for k, v in self.data_vars.items()})
        out.obs.operator = self.obs.operator
        return out
