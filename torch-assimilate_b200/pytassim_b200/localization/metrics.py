"""Distance functions understood by the B200 engine.

The reference takes an arbitrary Python ``dist_func(grid_ind, obs_grid)`` (pytassim/localization/gaspari_cohn.py:60-69,
125).  A GPU engine needs a closed set, so these objects are *both*: a numpy callable with the reference's calling
convention — ``grid_ind`` is a row ``[t_unix, coord_0, ...]`` of ``_extract_state_information``
(interface/mixin_local.py:50-69), ``obs_grid`` the observation info as DataFrame / (M, 1+nc) array
(mixin_local.py:45-47) — so the very same object can drive the reference's ``GaspariCohn`` on the CPU, and a
descriptor (``metric_id``, ``params``, ``n_coord``) that selects the device implementation.
Foreign callables raise ``NotImplementedError`` in the engine (no CPU fallback).
"""
import numpy as np

from .. import _cabi

__all__ = ["AbsDistance1D", "PeriodicDistance1D", "EuclideanDistance", "HaversineDistance", "ZeroDistance",
           "ProductDistance"]


def _rows(obs_grid):
    values = getattr(obs_grid, "values", obs_grid)          # pandas.DataFrame or ndarray
    return np.asarray(values, dtype=np.float64)


class _Metric(object):
    metric_id = None
    params = ()
    n_coord = 1

    def __repr__(self):
        return "{0}({1})".format(type(self).__name__, ", ".join(str(p) for p in self.params))

    def cache_key(self):
        """Everything that distinguishes two metric objects for the engine (plans are cached per key)."""
        return (type(self).__name__, int(self.metric_id), int(self.n_coord), int(getattr(self, "n_extra", 0)),
                tuple(float(p) for p in self.params), bool(getattr(self, "zero_coords", False)))


class AbsDistance1D(_Metric):
    """|x_g - x_o| on one coordinate (examples/benchmark_letkf.py:85-87, testing/dummy.py:142-151)."""
    metric_id = _cabi.METRIC_ABS1D

    def __call__(self, grid_ind, obs_grid):
        obs = _rows(obs_grid)
        return np.abs(np.asarray(grid_ind, dtype=np.float64)[1] - obs[:, 1])


class ZeroDistance(_Metric):
    """Distance 0 between every grid point and every observation: every observation is local with weight 1, the LETKF
    degenerates to the global ETKF at every grid point.  This is the ``dist_func=lambda x, y: np.zeros(y.shape[0])`` of the
    reference's own tests (tests/unit_tests/interface/test_letkf.py:79-104); on the device it is |x_g - x_o| with every
    coordinate replaced by 0 (``zero_coords``)."""
    metric_id = _cabi.METRIC_ABS1D
    zero_coords = True

    def __call__(self, grid_ind, obs_grid):
        return np.zeros(_rows(obs_grid).shape[0])


class PeriodicDistance1D(_Metric):
    """min(|x_g - x_o|, L - |x_g - x_o|) on a ring of circumference L (Lorenz-96 style); coordinates in [0, L]."""
    metric_id = _cabi.METRIC_PERIODIC1D

    def __init__(self, period):
        self.period = float(period)
        self.params = (self.period,)

    def __call__(self, grid_ind, obs_grid):
        obs = _rows(obs_grid)
        d = np.abs(np.asarray(grid_ind, dtype=np.float64)[1] - obs[:, 1])
        return np.minimum(d, self.period - d)


class EuclideanDistance(_Metric):
    """sqrt(sum_c (x_gc - x_oc)^2) over 1..3 coordinate columns."""
    metric_id = _cabi.METRIC_EUCLID

    def __init__(self, n_coord=2):
        if n_coord not in (1, 2, 3):
            raise ValueError("EuclideanDistance supports 1 to 3 coordinates")
        self.n_coord = int(n_coord)

    def __call__(self, grid_ind, obs_grid):
        obs = _rows(obs_grid)
        diff = obs[:, 1:1 + self.n_coord] - np.asarray(grid_ind, dtype=np.float64)[None, 1:1 + self.n_coord]
        acc = diff[:, 0] * diff[:, 0]
        for c in range(1, self.n_coord):
            acc = acc + diff[:, c] * diff[:, c]
        return np.sqrt(acc)


class HaversineDistance(_Metric):
    """Great-circle distance on a sphere of the given radius; coordinate columns are (lat, lon) in degrees."""
    metric_id = _cabi.METRIC_HAVERSINE
    n_coord = 2

    def __init__(self, radius=6371.0):
        self.radius = float(radius)
        self.params = (self.radius,)

    def __call__(self, grid_ind, obs_grid):
        obs = _rows(obs_grid)
        g = np.asarray(grid_ind, dtype=np.float64)
        phi1, lam1 = np.radians(g[1]), np.radians(g[2])
        phi2, lam2 = np.radians(obs[:, 1]), np.radians(obs[:, 2])
        a = np.sin((phi2 - phi1) / 2) ** 2 + np.cos(phi1) * np.cos(phi2) * np.sin((lam2 - lam1) / 2) ** 2
        a = np.clip(a, 0.0, 1.0)
        return 2.0 * self.radius * np.arcsin(np.sqrt(a))


class ProductDistance(_Metric):
    """Several distance rows: row 0 is ``primary`` on the first ``primary.n_coord`` coordinate columns, rows 1.. are
    ``|x_g - x_o|`` on the following ``n_extra`` coordinate columns (vertical level, ...; at most 2).  ``GaspariCohn`` takes
    one ``length_scale`` entry per row and multiplies the tapers (pytassim/localization/gaspari_cohn.py:124-134).  Columns are
    the coordinate levels AFTER the time column of the grid / observation rows (interface/mixin_local.py:45-69): the time
    column itself is not handed to the device (``LETKF._analyse_arrays`` drops it), so a temporal taper needs the time offset
    as an explicit coordinate level of the grid index and the observation info."""

    def __init__(self, primary, n_extra=1):
        if not isinstance(primary, _Metric) or isinstance(primary, ProductDistance) or getattr(primary, "zero_coords", False):
            raise ValueError("primary must be one of the single-row metric objects")
        if n_extra not in (1, 2):
            raise ValueError("ProductDistance supports 1 or 2 extra components")
        self.primary = primary
        self.n_extra = int(n_extra)
        self.metric_id = primary.metric_id
        self.params = primary.params
        self.n_coord = primary.n_coord + self.n_extra

    def __repr__(self):
        return "ProductDistance({0!r}, n_extra={1})".format(self.primary, self.n_extra)

    def __call__(self, grid_ind, obs_grid):
        obs = _rows(obs_grid)
        g = np.asarray(grid_ind, dtype=np.float64)
        np_ = self.primary.n_coord
        rows = [self.primary(g[:1 + np_], obs[:, :1 + np_])]
        for e in range(self.n_extra):
            rows.append(np.abs(g[1 + np_ + e] - obs[:, 1 + np_ + e]))
        return np.stack(rows, axis=0)
