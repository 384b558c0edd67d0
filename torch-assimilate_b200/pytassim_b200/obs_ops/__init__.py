"""Observation operators that SELECT grid columns of one state variable, usable as ``ds.obs.operator``.

Reference: pytassim/obs_ops/base_ops.py:42-87 (``BaseOperator.__call__``: apply ``obs_op``, pick the observation times,
rename ``grid`` -> ``obs_grid_1``), pytassim/obs_ops/lorenz_96/identity.py:41-94 (``IdentityOperator``: label selection of
prescribed / random / all grid points of variable ``x``) and examples/benchmark_letkf.py:90-104 (nearest-grid-point
selection at evenly spaced positions).

Called like the reference's operators (``operator(obs_ds, state)``) they return the observation equivalents as a host array
with dims ``('time', 'ensemble', 'obs_grid_1')``.  They additionally expose ``device_index(obs_ds, state)``: the (variable,
time, grid column) positions of the selection.  ``FilterAssimilation.update_state`` uses it to gather the ensemble of
observation equivalents from the state that is already on the GPU, fused with the ensemble mean / perturbation / innovation /
R^-1/2 step (``b200da_obs_gather_prep``) instead of building a (k, M) host array and uploading it.
"""
import numpy as np
import pandas as pd

from ..xrlite import DataArray

__all__ = ["GridSelectOperator", "IdentityOperator", "NearestGridOperator", "PositionOperator"]


class GridSelectOperator(object):
    """Base class: ``grid_positions(grid_index) -> int array`` defines the selected grid columns."""

    def __init__(self, len_grid=40, random_state=None, var_name='x'):
        self.len_grid = len_grid                      # base_ops.py:60-61
        self.random_state = random_state
        self.var_name = var_name

    def grid_positions(self, grid_index):
        raise NotImplementedError

    def device_index(self, obs_ds, state):
        """(var position, time positions (n_time_obs,), grid positions (n_obs,)) inside ``state`` (var_name, time, ensemble,
        grid).  KeyError for observation times or grid labels that the state does not have (xarray ``sel`` semantics)."""
        var_pos = 0
        if 'var_name' in state.dims:                                               # identity.py:89-90
            var_pos = pd.Index(state.indexes['var_name']).get_indexer([self.var_name])[0]
            if var_pos < 0:
                raise KeyError(self.var_name)
        obs_times = pd.Index(obs_ds['observations'].indexes['time'])
        t_pos = pd.Index(state.indexes['time']).get_indexer(obs_times)             # base_ops.py:70
        if (t_pos < 0).any():
            raise KeyError(obs_times[t_pos < 0][0])
        g_pos = np.asarray(self.grid_positions(state.indexes['grid']), dtype=np.int64)
        n_obs = len(obs_ds['observations'].indexes['obs_grid_1'])
        if g_pos.shape[0] != n_obs:                                                # base_ops.py:73 (assigning a wrong-sized index)
            raise ValueError('conflicting sizes for dimension obs_grid_1: operator selects {0:d} grid points, observations '
                             'have {1:d}'.format(g_pos.shape[0], n_obs))
        return int(var_pos), np.asarray(t_pos, dtype=np.int64), g_pos

    def __call__(self, obs_ds, input_vals, *args, **kwargs):
        """Host path with the reference's semantics (base_ops.py:63-75)."""
        var_pos, t_pos, g_pos = self.device_index(obs_ds, input_vals)
        vals = np.asarray(input_vals.values)
        if 'var_name' in input_vals.dims:
            vals = np.take(vals, var_pos, axis=input_vals.dims.index('var_name'))
        dims = [d for d in input_vals.dims if d != 'var_name']
        vals = np.take(vals, t_pos, axis=dims.index('time'))
        vals = np.take(vals, g_pos, axis=dims.index('grid'))
        coords = {d: input_vals.indexes[d] for d in dims if d not in ('time', 'grid')}
        coords['time'] = obs_ds['observations'].indexes['time']
        coords['obs_grid_1'] = obs_ds['observations'].indexes['obs_grid_1']
        return DataArray(vals, coords, ['obs_grid_1' if d == 'grid' else d for d in dims])


class IdentityOperator(GridSelectOperator):
    """identity.py:41-94: observed grid points equal observations.  ``obs_points``: int -> that many points drawn without
    replacement from ``random_state``; list -> prescribed grid labels; None -> every grid point."""

    def __init__(self, obs_points=None, len_grid=40, random_state=None, var_name='x'):
        super().__init__(len_grid=len_grid, random_state=random_state, var_name=var_name)
        self._obs_points = None
        self._sel_obs_points = None
        self.obs_points = obs_points

    @property
    def obs_points(self):
        return self._obs_points

    @obs_points.setter
    def obs_points(self, points):                                                  # identity.py:74-86
        if isinstance(points, (int, float)):
            self._sel_obs_points = self.random_state.choice(self.len_grid, size=points, replace=False)
        elif points is None:
            self._sel_obs_points = np.arange(self.len_grid)
        else:
            self._sel_obs_points = points
        self._obs_points = points

    def grid_positions(self, grid_index):                                          # identity.py:91: sel(grid=labels)
        index = pd.Index(grid_index)
        pos = index.get_indexer(pd.Index(self._sel_obs_points) if not isinstance(index, pd.MultiIndex)
                                else list(self._sel_obs_points))
        if (pos < 0).any():
            raise KeyError(np.asarray(self._sel_obs_points, dtype=object)[pos < 0][0])
        return pos


class NearestGridOperator(GridSelectOperator):
    """examples/benchmark_letkf.py:90-104: ``nr_obs`` evenly spaced positions ``linspace(0, len_grid, nr_obs, endpoint=False)``
    (or explicit ``obs_grid`` positions) observed at the nearest grid point (pandas nearest-label semantics, ties to the larger
    label, exactly what ``DataArray.sel(method='nearest')`` resolves to)."""

    def __init__(self, len_grid=40, nr_obs=None, obs_grid=None, var_name='x'):
        super().__init__(len_grid=len_grid, var_name=var_name)
        self.nr_obs = nr_obs
        self._obs_grid = None if obs_grid is None else np.asarray(obs_grid, dtype=np.float64)

    @property
    def obs_grid(self):                                                            # benchmark_letkf.py:95-98
        if self._obs_grid is not None:
            return self._obs_grid
        return np.linspace(start=0, stop=self.len_grid, num=self.nr_obs, endpoint=False)

    def grid_positions(self, grid_index):
        return pd.Index(grid_index).get_indexer(self.obs_grid, method='nearest')


class PositionOperator(GridSelectOperator):
    """Selection by precomputed integer grid POSITIONS (not labels), e.g. the nearest-neighbour indices a KD-tree lookup
    returns (the pattern of obs_ops/terrsysmp/cos_t2m.py:137-143 without its lapse-rate correction)."""

    def __init__(self, positions, var_name='x'):
        positions = np.asarray(positions, dtype=np.int64)
        super().__init__(len_grid=None, var_name=var_name)
        self.positions = positions

    def grid_positions(self, grid_index):
        if self.positions.size and (self.positions.min() < 0 or self.positions.max() >= len(grid_index)):
            raise IndexError("grid position outside the state grid")
        return self.positions
