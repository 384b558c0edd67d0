"""LETKF on the B200 engine.  Reference: pytassim/interface/letkf.py:38-148, interface/mixin_local.py:33-69."""
import numpy as np
import torch

from .base import index_to_array, to_host
from .etkf import ETKF
from ..engine import LETKFEngine

__all__ = ['LETKF']


class LETKF(ETKF):
    """Same constructor as the reference (letkf.py:72-92).  ``gpu`` and ``chunksize`` are accepted and ignored: the
    analysis always runs on the GPU and grid points are grouped into CTA blocks by the engine, not by dask chunks."""

    def __init__(self, localization=None, inf_factor=1.0, smoother=False, gpu=False, pre_transform=None,
                 post_transform=None, chunksize=10, weight_save_path=None, forward_model=None):
        super().__init__(inf_factor=inf_factor, smoother=smoother, gpu=gpu, pre_transform=pre_transform,
                         post_transform=post_transform, weight_save_path=weight_save_path, forward_model=forward_model)
        self.localization = localization
        self.chunksize = chunksize
        self._grid_cache = None

    def __str__(self):
        return 'Localized ETKF(inf_factor={0}, loc={1})'.format(str(self.inf_factor.item()), str(self.localization))

    def __repr__(self):
        return 'LETKF({0},{1})'.format(repr(self.inf_factor.item()), repr(self.localization))

    @property
    def chunks(self):
        return dict(grid=self.chunksize)                   # mixin_local.py:34-36

    @property
    def localized_module(self):
        """mixin_local.py:37-42: one grid point per call (numpy in / out) — the per-grid-point form of the hot loop."""
        from .per_point import LocalObservations
        return LocalObservations(module=self.module, localization=self.localization)

    def _local_engine(self, k, n_slices, grid_coords):
        loc = self.localization
        key = ('local', k, n_slices, float(self.inf_factor), self.dtype, type(loc).__name__, loc.dist_func.cache_key(),
               tuple(np.atleast_1d(loc.radius).tolist()), float(loc.epsilon), self._kernel_key())
        if key not in self._engines:
            self._engines = {kk: v for kk, v in self._engines.items() if kk[0] != 'local'}
            self._engines[key] = self._configure_engine(
                LETKFEngine(k, n_slices, loc.dist_func, loc.radius, epsilon=loc.epsilon, inf_factor=float(self.inf_factor),
                            taper=loc.taper, dtype=self.dtype))
            self._grid_cache = None
        eng = self._engines[key]
        if self._grid_cache is None or self._grid_cache.shape != grid_coords.shape or \
                not np.array_equal(self._grid_cache, grid_coords):
            eng.set_grid(grid_coords)
            self._grid_cache = grid_coords.copy()
        return eng

    def _analyse_arrays(self, state, x, innov, perts, obs_info):
        if self.localization is None:                      # letkf.py / wrapper.py:87: plain ETKF for every grid point
            return super()._analyse_arrays(state, x, innov, perts, obs_info)
        nc = self.localization.dist_func.n_coord
        if getattr(self.localization.dist_func, 'zero_coords', False):      # ZeroDistance: every pair at distance 0
            grid_coords = np.zeros((len(state.indexes['grid']), nc))
            obs_coords = np.zeros((obs_info.shape[0], nc))
        else:
            grid_coords = index_to_array(state.indexes['grid'])             # mixin_local.py:50-69
            if grid_coords.shape[1] < nc or obs_info.shape[1] - 1 < nc:
                raise ValueError("the metric needs {0} coordinate column(s)".format(nc))
            obs_coords = obs_info[:, 1:1 + nc]
        eng = self._local_engine(x.shape[1], x.shape[0], np.ascontiguousarray(grid_coords[:, :nc]))
        eng.bin_obs(obs_coords, perts, innov)
        if self.weight_save_path is not None:
            # filter.py:159-162: the weights of every grid point leave the device, go through the netCDF store, come back and
            # are applied by the stand-alone update kernel (b200da_apply_weights) instead of the fused tail of the solve
            xd = torch.as_tensor(x).to(eng.device)
            _, weights = eng.analyse(xd, return_weights=True)
            weights = self._weights_through_store(state, weights.cpu().numpy())
            return to_host(eng.apply_weights(xd, weights))
        return to_host(eng.analyse(torch.as_tensor(x)))
