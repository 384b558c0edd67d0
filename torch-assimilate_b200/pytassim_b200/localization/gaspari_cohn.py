"""Gaspari-Cohn localization descriptors with the reference's constructor signature
(pytassim/localization/gaspari_cohn.py:60-69, :153-162).

The objects carry no numerics of their own: ``LETKF`` hands (taper, length scale, epsilon, metric) to the CUDA
engine, where the taper is fused into the staging of the local observations.  ``localize_obs`` keeps the reference
signature (localization/localization.py:53-80) and answers from the GPU neighbour search.  ``dist_func`` must be one
of the metric objects of :mod:`pytassim_b200.localization.metrics`; an arbitrary Python callable cannot run on the
device and raises ``NotImplementedError`` (there is no CPU fallback).
"""
import numpy as np

from .metrics import _Metric

__all__ = ["GaspariCohn", "GaspariCohnInf", "BaseLocalization"]


class BaseLocalization(object):
    taper = None

    def _engine_for(self, ens_size=2):
        from ..engine import LETKFEngine
        return LETKFEngine(ens_size, 1, self.dist_func, self.radius, epsilon=self.epsilon, taper=self.taper)

    def localize_obs(self, grid_ind, obs_grid):
        """(use_obs bool[M], weights float[M]) for one grid row ``[t, coords...]`` and the (M, 1+nc) observation
        info; weights are returned for the used observations, 0 elsewhere."""
        obs = np.asarray(getattr(obs_grid, "values", obs_grid), dtype=np.float64)
        grid = np.asarray(grid_ind, dtype=np.float64).reshape(1, -1)
        nc = self.dist_func.n_coord
        if getattr(self.dist_func, 'zero_coords', False):
            obs = np.zeros_like(obs[:, :1 + nc]); grid = np.zeros_like(grid[:, :1 + nc])
        eng = self._engine_for()
        eng.set_grid(grid[:, 1:1 + nc])
        m = obs.shape[0]
        eng.bin_obs(obs[:, 1:1 + nc], np.zeros((2, m)), np.zeros(m))
        off, idx, w, _, _ = eng.neighbour_lists()
        use = np.zeros(m, dtype=bool)
        weights = np.zeros(m, dtype=np.float64)
        sel = idx.cpu().numpy()
        use[sel] = True
        weights[sel] = w.cpu().numpy()
        return use, weights


def _check_metric(dist_func):
    if not isinstance(dist_func, _Metric):
        raise NotImplementedError(
            "the B200 engine evaluates distances on the device: dist_func must be a pytassim_b200.localization.metrics "
            "object (AbsDistance1D, PeriodicDistance1D, EuclideanDistance, HaversineDistance), got {0!r}".format(dist_func))


class GaspariCohn(BaseLocalization):
    taper = "gc"

    def __init__(self, length_scale, dist_func, epsilon=1E-5):
        _check_metric(dist_func)
        self.radius = np.atleast_1d(length_scale)
        self.dist_func = dist_func
        self.epsilon = epsilon

    def __str__(self):
        return 'GaspariCohn(l={0})'.format(str(self.radius))

    def __repr__(self):
        return 'GaspariCohn'


class GaspariCohnInf(BaseLocalization):
    taper = "gcinf"

    def __init__(self, length_scale, dist_func, epsilon=1E-5):
        _check_metric(dist_func)
        self.radius = np.atleast_1d(length_scale)
        self.dist_func = dist_func
        self.epsilon = epsilon

    def __str__(self):
        return 'GaspariCohnInf(l={0})'.format(str(self.radius))

    def __repr__(self):
        return 'GaspariCohnInf'
