"""Gaspari-Cohn localization descriptors with the reference's constructor signature
(pytassim/localization/gaspari_cohn.py:60-69, :153-162).

``LETKF`` hands (taper, length scale, epsilon, metric) to the CUDA engine, where the taper is fused into the staging of
the local observations.  ``localize_obs`` keeps the reference signature (localization/localization.py:53-80) and answers
from the GPU neighbour search.  The only numerics on the host are ``host_decision``: the reference's own numpy expression
for the handful of (grid point, observation) pairs per analysis whose taper value the device finds within 1e-13 of
``epsilon`` — there the mask ``weights > epsilon`` (gaspari_cohn.py:135) depends on the last bits of numpy's ``**`` on
this host, so the host decides and the device applies the decision (the ambiguity protocol of include/b200da.h).  ``dist_func`` must be one
of the metric objects of :mod:`pytassim_b200.localization.metrics`; an arbitrary Python callable cannot run on the
device and raises ``NotImplementedError`` (there is no CPU fallback).
"""
import numpy as np

from .metrics import _Metric

__all__ = ["GaspariCohn", "GaspariCohnInf", "BaseLocalization"]


def _gc_inner(r):                     # gaspari_cohn.py:78-84, same operations in the same order
    f = - 0.25 * r ** 5
    f += 0.5 * r ** 4
    f += 0.625 * r ** 3
    f -= 5 / 3 * r ** 2
    f += 1
    return f


def _gc_outer(r):                     # gaspari_cohn.py:87-95
    f = 1 / 12 * r ** 5
    f -= 0.5 * r ** 4
    f += 0.625 * r ** 3
    f += 5 / 3 * r ** 2
    f -= 5 * r
    f += 4
    f -= 2 / 3 / r
    return f


def _gcinf_pieces(r):                 # gaspari_cohn.py:172-210, pieces for r < 0.5, 1, 1.5, 2
    f1 = -28 * r ** 5 / 33
    f1 += 8 * r ** 4 / 11
    f1 += 20 * r ** 3 / 11
    f1 -= 80 * r ** 2 / 33
    f1 += 1
    f2 = 20 * r ** 5 / 33
    f2 -= 16 * r ** 4 / 11
    f2 += 100 * r ** 2 / 33
    f2 -= 45 * r / 11
    f2 += 51 / 22
    f2 -= 7 / (44 * r)
    f3 = -4 * r ** 5 / 11
    f3 += 16 * r ** 4 / 11
    f3 -= 10 * r ** 3 / 11
    f3 -= 100 * r ** 2 / 33
    f3 += 5 * r
    f3 -= 61 / 22
    f3 += 115 / (132 * r)
    f4 = 4 * r ** 5 / 33
    f4 -= 8 * r ** 4 / 11
    f4 += 10 * r ** 3 / 11
    f4 += 80 * r ** 2 / 33
    f4 -= 80 * r / 11
    f4 += 64 / 11
    f4 -= 32 / (33 * r)
    return f1, f2, f3, f4


class BaseLocalization(object):
    taper = None

    def host_decision(self, grid_row, obs_rows):
        """(use bool[m], weights float[m]) for one grid row ``[t, coords...]`` and a FEW observation rows, by the
        reference's numpy expression (gaspari_cohn.py:120-135 / :240-254) on this host.  Only called for pairs the device
        reports as ambiguous (taper value within 1e-13 of epsilon); the bulk of the taper never runs on the host."""
        import warnings
        obs_rows = np.asarray(obs_rows, dtype=np.float64)
        m = obs_rows.shape[0]
        if getattr(self.dist_func, 'zero_coords', False):
            dist = np.zeros((1, m))
        else:
            dist = np.atleast_2d(self.dist_func(np.asarray(grid_row, dtype=np.float64), obs_rows))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if self.taper == "gcinf":
                r = dist[0] / np.atleast_1d(self.radius)[0]
                weights = np.zeros(m, dtype=float)
                f1, f2, f3, f4 = _gcinf_pieces(r)
                for cond, f in ((r < 2, f4), (r < 1.5, f3), (r < 1, f2), (r < 0.5, f1)):
                    weights[cond] = f[cond]
            else:
                weights = np.ones(m, dtype=float)
                for i, d in enumerate(dist):
                    r = d / self.radius[i]
                    tmp = np.zeros(m, dtype=float)
                    tmp[r < 2] = _gc_outer(r[r < 2])
                    tmp[r < 1] = _gc_inner(r[r < 1])
                    weights *= tmp
        return weights > self.epsilon, weights

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
