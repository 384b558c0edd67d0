"""CPU oracle for the LETKF / ETKF analysis hot path of tobifinn/torch-assimilate (pytassim 0.2.1).

TEST INFRASTRUCTURE ONLY.  This module is the *checker*: it may be imported by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` and by
nothing else.  The product (``torch-assimilate_b200/``) never imports it and has no CPU fallback.

It is a numpy restatement of the reference's per-grid-point algorithm; every function cites the
reference ``file:line`` it follows (paths relative to ``/root/reference``).  The reference is 100 %
Python, so there is no C to compile: the restatement is numpy + LAPACK (``numpy.linalg.eigh``), which
is what the reference itself resolves to (``torch.symeig`` -> LAPACK ``syevd``).

Parity status: **pinned**.  ``oracle/make_golden.py`` imports the reference's own leaf modules
(``pytassim/core/{base,utils,etkf}.py``, ``pytassim/localization/*.py``, ``pytassim/interface/wrapper.py``)
from ``/root/reference`` by file path (one shim: ``torch.symeig -> torch.linalg.eigh``, the API was
removed from torch), runs them on the reference's fixtures ``tests/data/test_state.nc`` /
``test_single_obs.nc`` and on seeded synthetic inputs, and commits the outputs under ``tests/golden/``.
``tests/test_oracle.py`` checks this restatement against those vectors and against the known-answer
values in the reference's unit tests (``tests/unit_tests/core/test_etkf.py:47-240``,
``tests/unit_tests/localization/test_gaspari_cohn.py:52-171``).  The xarray glue
(``interface/base.py:223-241,257-278,359-379``) cannot be imported here (xarray/dask are absent), so
it is restated and pinned through the reference test ``interface/test_letkf.py:106-157`` whose
hand-loop is reproduced with the imported leaves.
"""
from __future__ import annotations

import warnings
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

EARTH_RADIUS_KM = 6371.0

# ----------------------------------------------------------------------------------------------
# Localization  (pytassim/localization/gaspari_cohn.py)
# ----------------------------------------------------------------------------------------------


def gc_f1(dist: np.ndarray) -> np.ndarray:
    """Gaspari-Cohn inner polynomial, r < 1.  Reference: localization/gaspari_cohn.py:78-84
    (same expression order, ``**`` operators included, so it is bit-equal on the same host)."""
    f1 = - 0.25 * dist ** 5
    f1 += 0.5 * dist ** 4
    f1 += 0.625 * dist ** 3
    f1 -= 5 / 3 * dist ** 2
    f1 += 1
    return f1


def gc_f2(dist: np.ndarray) -> np.ndarray:
    """Gaspari-Cohn outer polynomial, 1 <= r < 2.  Reference: localization/gaspari_cohn.py:87-95."""
    f2 = 1 / 12 * dist ** 5
    f2 -= 0.5 * dist ** 4
    f2 += 0.625 * dist ** 3
    f2 += 5 / 3 * dist ** 2
    f2 -= 5 * dist
    f2 += 4
    f2 -= 2 / 3 / dist
    return f2


def gaspari_cohn_localize(dist, radius, epsilon: float = 1e-5) -> Tuple[np.ndarray, np.ndarray]:
    """``GaspariCohn.localize_obs`` given the already evaluated distance component(s).

    Reference: localization/gaspari_cohn.py:97-136.  ``dist`` is what ``dist_func(grid_ind, obs_grid)``
    returned: one array (M,) or a tuple / 2-D array of components; component *i* is divided by
    ``radius[i]``; f2 is written where r < 2, then f1 overwrites where r < 1 (so the -inf of f2 at r = 0
    never survives); the component tapers are multiplied; ``use = weights > epsilon``.
    """
    radius = np.atleast_1d(radius)                      # gaspari_cohn.py:66
    dist = np.atleast_2d(dist)                          # :125
    n_obs = dist.shape[-1]
    weights = np.ones((n_obs,), dtype=float)            # :124
    for i, d in enumerate(dist):                        # :126
        dist_radius = d / radius[i]                     # :127
        conds = [dist_radius < thres for thres in (2, 1)]   # :128, thresholds :69
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with np.errstate(all="ignore"):
                tmp_weights = np.zeros((n_obs,), dtype=float)
                tmp_weights[conds[0]] = gc_f2(dist_radius[conds[0]])     # :132
                tmp_weights[conds[1]] = gc_f1(dist_radius[conds[1]])     # :133
        weights *= tmp_weights                          # :134
    use_obs = weights > epsilon                         # :135
    return use_obs, weights


def gcinf_f1(dist):
    """Reference: localization/gaspari_cohn.py:172-178."""
    f1 = -28 * dist ** 5 / 33
    f1 += 8 * dist ** 4 / 11
    f1 += 20 * dist ** 3 / 11
    f1 -= 80 * dist ** 2 / 33
    f1 += 1
    return f1


def gcinf_f2(dist):
    """Reference: localization/gaspari_cohn.py:181-188."""
    f2 = 20 * dist ** 5 / 33
    f2 -= 16 * dist ** 4 / 11
    f2 += 100 * dist ** 2 / 33
    f2 -= 45 * dist / 11
    f2 += 51 / 22
    f2 -= 7 / (44 * dist)
    return f2


def gcinf_f3(dist):
    """Reference: localization/gaspari_cohn.py:191-199."""
    f3 = -4 * dist ** 5 / 11
    f3 += 16 * dist ** 4 / 11
    f3 -= 10 * dist ** 3 / 11
    f3 -= 100 * dist ** 2 / 33
    f3 += 5 * dist
    f3 -= 61 / 22
    f3 += 115 / (132 * dist)
    return f3


def gcinf_f4(dist):
    """Reference: localization/gaspari_cohn.py:202-210."""
    f4 = 4 * dist ** 5 / 33
    f4 -= 8 * dist ** 4 / 11
    f4 += 10 * dist ** 3 / 11
    f4 += 80 * dist ** 2 / 33
    f4 -= 80 * dist / 11
    f4 += 64 / 11
    f4 -= 32 / (33 * dist)
    return f4


def gaspari_cohn_inf_localize(dist, radius: float, epsilon: float = 1e-5):
    """``GaspariCohnInf.localize_obs`` given the evaluated distance (single component, scalar radius).

    Reference: localization/gaspari_cohn.py:216-254 (4-piece taper, thresholds 2, 1.5, 1, 0.5)."""
    dist = np.asarray(dist, dtype=float)
    weights = np.zeros((dist.shape[-1],), dtype=float)      # :244
    dist_radius = dist / radius                             # :246
    conds = [dist_radius < thres for thres in (2, 1.5, 1, 0.5)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            weights[conds[0]] = gcinf_f4(dist_radius[conds[0]])
            weights[conds[1]] = gcinf_f3(dist_radius[conds[1]])
            weights[conds[2]] = gcinf_f2(dist_radius[conds[2]])
    weights[conds[3]] = gcinf_f1(dist_radius[conds[3]])
    use_obs = weights > epsilon
    return use_obs, weights


# ----------------------------------------------------------------------------------------------
# Distance functions.  The reference leaves ``dist_func`` to the user (gaspari_cohn.py:60-69); these
# are the closed set the GPU engine implements, written the way a user of the reference would write
# them (cf. examples/benchmark_letkf.py:85-87, pytassim/testing/dummy.py:142-151,
# tests/unit_tests/interface/test_letkf.py:107-110).  ``grid_row`` is one row of
# ``_extract_state_information`` = [t_unix, coord_0, ...] (interface/mixin_local.py:50-69);
# ``obs_rows`` is the (M, 1+nc) array [time, coord_0, ...] (mixin_local.py:45-47).
# ----------------------------------------------------------------------------------------------


def dist_abs1d(grid_row, obs_rows):
    return np.abs(grid_row[1] - obs_rows[:, 1])


def make_dist_periodic1d(period: float) -> Callable:
    def dist_periodic1d(grid_row, obs_rows):
        d = np.abs(grid_row[1] - obs_rows[:, 1])
        return np.minimum(d, period - d)
    return dist_periodic1d


def dist_euclid(grid_row, obs_rows):
    diff = obs_rows[:, 1:] - np.asarray(grid_row)[None, 1:]
    acc = diff[:, 0] * diff[:, 0]
    for c in range(1, diff.shape[1]):
        acc = acc + diff[:, c] * diff[:, c]
    return np.sqrt(acc)


def make_dist_haversine(radius_km: float = EARTH_RADIUS_KM) -> Callable:
    """Great-circle distance; coordinates are (lat, lon) in degrees."""
    def dist_haversine(grid_row, obs_rows):
        phi1 = np.radians(grid_row[1])
        lam1 = np.radians(grid_row[2])
        phi2 = np.radians(obs_rows[:, 1])
        lam2 = np.radians(obs_rows[:, 2])
        a = np.sin((phi2 - phi1) / 2) ** 2 + np.cos(phi1) * np.cos(phi2) * np.sin((lam2 - lam1) / 2) ** 2
        a = np.clip(a, 0.0, 1.0)
        return 2.0 * radius_km * np.arcsin(np.sqrt(a))
    return dist_haversine


def make_dist_product(primary: Callable, n_primary: int, n_extra: int) -> Callable:
    """A dist_func returning SEVERAL rows (gaspari_cohn.py:125 ``np.atleast_2d``): row 0 = ``primary`` on the first
    ``n_primary`` coordinate columns, rows 1.. = |x_g - x_o| on the following ``n_extra`` columns."""
    def dist_product(grid_row, obs_rows):
        obs_rows = np.asarray(obs_rows)
        rows = [primary(grid_row[:1 + n_primary], obs_rows[:, :1 + n_primary])]
        for e in range(n_extra):
            rows.append(np.abs(grid_row[1 + n_primary + e] - obs_rows[:, 1 + n_primary + e]))
        return np.stack(rows, axis=0)
    return dist_product


# ----------------------------------------------------------------------------------------------
# Core ETKF weights  (pytassim/core/{base,utils,etkf}.py)
# ----------------------------------------------------------------------------------------------


def evd(tensor: np.ndarray, reg_value: float = 0.0):
    """Reference: core/utils.py:26-61 — symeig(lower), clamp(min=0), + reg, reciprocal."""
    evals, evects = np.linalg.eigh(tensor, UPLO="L")     # utils.py:57
    evals = np.clip(evals, 0, None)                      # :58
    evals = evals + reg_value                            # :59
    evals_inv = 1 / evals                                # :60
    return evals, evects, evals_inv


def rev_evd(evals: np.ndarray, evects: np.ndarray) -> np.ndarray:
    """Reference: core/utils.py:64-93 — U diag(evals) U^T as two matrix products."""
    rev_mat = evects @ np.diag(evals)                    # :90-91
    rev_mat = rev_mat @ evects.T                         # :92
    return rev_mat


def etkf_estimate_weights(normed_perts: np.ndarray, normed_obs: np.ndarray, inf_factor: float):
    """Reference: core/etkf.py:57-77.  normed_perts (k, p), normed_obs (1, p).
    Returns (w_mean (k,1), w_perts (k,k), cov_analysed (k,k))."""
    ens_size = normed_perts.shape[-2]
    reg_value = (ens_size - 1) / inf_factor                       # etkf.py:67
    kernel_perts = normed_perts @ normed_perts.T                  # :68, utils.py:172
    evals, evects, evals_inv = evd(kernel_perts, reg_value)       # :69
    cov_analysed = rev_evd(evals_inv, evects)                     # :70
    kernel_obs = normed_perts @ normed_obs.T                      # :72
    w_mean = cov_analysed @ kernel_obs                            # :73
    square_root_einv = np.sqrt((ens_size - 1) * evals_inv)        # :75
    w_perts = rev_evd(square_root_einv, evects)                   # :76
    return w_mean, w_perts, cov_analysed


def etkf_weights(normed_perts: np.ndarray, normed_obs: np.ndarray, inf_factor: float = 1.0) -> np.ndarray:
    """``ETKFModule.forward``.  Reference: core/etkf.py:79-103, core/base.py:28-62.

    W[i, j] = w_mean[i] + w_perts[i, j]; empty observations -> sqrt(inf_factor) * I; a size mismatch
    between the last dims raises ValueError with the reference's message."""
    normed_perts = np.asarray(normed_perts, dtype=float)
    normed_obs = np.asarray(normed_obs, dtype=float)
    if normed_perts.shape[-1] != normed_obs.shape[-1]:            # base.py:33-38
        raise ValueError(
            'Observational size between ensemble ({0:d}) and observations '
            '({1:d}) do not match!'.format(normed_perts.shape[-1], normed_obs.shape[-1])
        )
    ens_size = normed_perts.shape[-2]
    if normed_perts.shape[-1] == 0:                               # etkf.py:91-95
        w_mean = np.zeros((ens_size, 1))                          # base.py:56
        w_perts = np.eye(ens_size) * np.sqrt(inf_factor)          # base.py:57-60, etkf.py:95
    else:
        normed_perts = normed_perts.reshape(-1, normed_perts.shape[-1])    # base.py:41-46
        normed_obs = normed_obs.reshape(1, -1) if normed_obs.ndim < 2 \
            else normed_obs.reshape(-1, normed_obs.shape[-1])
        w_mean, w_perts, _ = etkf_estimate_weights(normed_perts, normed_obs, inf_factor)
    return w_mean + w_perts                                       # etkf.py:102


class OracleKernel(object):
    """A kernel of pytassim/kernels as a plain numpy function ``fn(x (n, p), y (m, p)) -> (n, m)`` with the three
    compositions of kernels/base_kernels.py:40-161 (``+`` :89-90, ``*`` :119-120, ``**`` :160-161)."""

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, x, y):
        return self.fn(np.asarray(x, dtype=float), np.asarray(y, dtype=float))

    def __add__(self, other):
        return OracleKernel(lambda x, y: self(x, y) + other(x, y))

    def __mul__(self, other):
        return OracleKernel(lambda x, y: self(x, y) * other(x, y))

    def __pow__(self, other):
        return OracleKernel(lambda x, y: np.power(self(x, y), other(x, y)))


def _p(value):
    """Kernel parameters may be 0-dim torch tensors (the reference's defaults are float32 tensors); torch promotes them to
    the dtype of the float64 data, i.e. to float(value)."""
    return float(value)


def kernel_distance_matrix(x, y, norm=2.0):
    """kernels/utils.py:60-86: ``torch.cdist(x, y, p=norm)``."""
    diff = np.abs(x[:, None, :] - y[None, :, :])
    if norm == 1.0:
        return diff.sum(axis=-1)
    return np.sqrt((diff * diff).sum(axis=-1))


def kernel_euclidean_dist(x, y):
    """kernels/utils.py:89-110: squared 2-norm distance."""
    return kernel_distance_matrix(x, y, 2.0) ** 2


def LinearKernel():
    """kernels/linear.py:62-63 (dot product, kernels/utils.py:36-57)."""
    return OracleKernel(lambda x, y: x @ y.T)


def GaussKernel(lengthscale=1.):
    """kernels/rbf.py GaussKernel.forward: exp(-|x/l - y/l|^2 / 2)."""
    ls = _p(lengthscale)
    return OracleKernel(lambda x, y: np.exp(-(kernel_euclidean_dist(x / ls, y / ls) / 2.)))


def RBFKernel(gamma=0.5):
    """kernels/rbf.py RBFKernel: GaussKernel with lengthscale (0.5 / gamma) ** 0.5, evaluated in the parameter's own type."""
    return GaussKernel((0.5 / gamma) ** 0.5)


def PolyKernel(degree=2., const=1.):
    """kernels/polynomial.py forward."""
    deg, c = _p(degree), _p(const)
    return OracleKernel(lambda x, y: np.power(x @ y.T + c, deg))


def TanhKernel(coeff=1., const=0.):
    """kernels/tanh.py forward."""
    a, c = _p(coeff), _p(const)
    return OracleKernel(lambda x, y: np.tanh(a * (x @ y.T) + c))


def RationalKernel(lengthscale=1., weighting=1.):
    """kernels/rational.py forward."""
    ls, w = _p(lengthscale), _p(weighting)
    return OracleKernel(lambda x, y: np.power(1 + kernel_euclidean_dist(x / ls, y / ls) / (2 * w), -w))


def _p32(value):
    """scale.py:70-72 and diag.py:66-71 build their matrix from ``torch.ones`` of the DEFAULT dtype (float32) and multiply by
    the scaling before anything is cast to the data's dtype: the constant the reference really uses is float32-rounded."""
    return float(np.float32(float(value)))


def ScaleKernel(scaling=0.):
    """kernels/scale.py forward."""
    c = _p32(scaling)
    return OracleKernel(lambda x, y: np.ones((x.shape[0], y.shape[0])) * c)


def DiagKernel(scaling=0.):
    """kernels/diag.py forward: zeros for different sample counts, else scaling * I."""
    c = _p32(scaling)
    return OracleKernel(lambda x, y: np.zeros((x.shape[0], y.shape[0])) if x.shape[0] != y.shape[0]
                        else np.eye(x.shape[0]) * c)


def OrnsteinUhlenbeckKernel(lengthscale=1.):
    """kernels/orn_uhl.py forward (CPU restatement only: the device path rejects L1 kernels)."""
    ls = _p(lengthscale)
    return OracleKernel(lambda x, y: np.exp(-kernel_distance_matrix(x, y, 1.0) / ls))


def ketkf_estimate_weights(normed_perts: np.ndarray, normed_obs: np.ndarray, inf_factor: float, kernel=None):
    """Reference: core/ketkf.py:69-100.  normed_perts (k, p), normed_obs (1, p), kernel: an ``OracleKernel`` (default: linear)
    -> (w_mean (k,1), w_perts (k,k), cov_analysed (k,k))."""
    kernel = LinearKernel() if kernel is None else kernel
    ens_size = normed_perts.shape[0]
    reg_value = (ens_size - 1) / inf_factor                                           # ketkf.py:78
    k_perts = kernel(normed_perts, normed_perts)                                      # :80
    k_partial_mean = k_perts.mean(axis=-1, keepdims=True)                             # :81
    k_partial_mean = k_partial_mean - k_partial_mean.mean(axis=-2, keepdims=True)     # :82-83
    k_perts_centered = k_perts - k_perts.mean(axis=-2, keepdims=True) - k_partial_mean   # :84-85
    evals, evects, evals_inv = evd(k_perts_centered, reg_value)                       # :87
    cov_analysed = rev_evd(evals_inv, evects)                                         # :88
    k_obs = kernel(normed_perts, normed_obs)                                          # :90
    k_obs_centered = k_obs - k_obs.mean(axis=-2, keepdims=True)                       # :91
    k_obs_centered = k_obs_centered - k_partial_mean                                  # :92
    w_mean = cov_analysed @ k_obs_centered                                            # :93
    square_root_einv = np.sqrt((ens_size - 1) * evals_inv)                            # :95
    w_perts = rev_evd(square_root_einv, evects)                                       # :96
    return w_mean, w_perts, cov_analysed


def ketkf_weights(normed_perts: np.ndarray, normed_obs: np.ndarray, inf_factor: float = 1.0, kernel=None) -> np.ndarray:
    """``KETKFModule(kernel).forward`` = ``ETKFModule.forward`` (core/etkf.py:79-103) with the kernelised
    ``_estimate_weights``."""
    normed_perts = np.asarray(normed_perts, dtype=float)
    normed_obs = np.asarray(normed_obs, dtype=float)
    if normed_perts.shape[-1] != normed_obs.shape[-1]:
        raise ValueError('Observational size between ensemble ({0:d}) and observations '
                         '({1:d}) do not match!'.format(normed_perts.shape[-1], normed_obs.shape[-1]))
    ens_size = normed_perts.shape[-2]
    if normed_perts.shape[-1] == 0:
        return np.eye(ens_size) * np.sqrt(inf_factor)
    w_mean, w_perts, _ = ketkf_estimate_weights(normed_perts.reshape(-1, normed_perts.shape[-1]),
                                                normed_obs.reshape(1, -1), inf_factor, kernel)
    return w_mean + w_perts


def ketkf_linear_estimate_weights(normed_perts: np.ndarray, normed_obs: np.ndarray, inf_factor: float):
    """core/ketkf.py:69-100 with ``LinearKernel`` (kernels/linear.py:62-63: K(x, y) = x y^T)."""
    return ketkf_estimate_weights(normed_perts, normed_obs, inf_factor, LinearKernel())


def ketkf_linear_weights(normed_perts: np.ndarray, normed_obs: np.ndarray, inf_factor: float = 1.0) -> np.ndarray:
    """``KETKFModule(LinearKernel).forward``."""
    return ketkf_weights(normed_perts, normed_obs, inf_factor, LinearKernel())


def lketkf_weights_point(grid_row, normed_perts, normed_obs, obs_rows, dist_func, radius, kernel,
                         epsilon=1e-5, inf_factor=1.0):
    """One grid point of the localized KETKF (interface/lketkf.py:84-110 -> wrapper.py:86-98 around ``KETKFModule``)."""
    luse, lweights = gaspari_cohn_localize(dist_func(grid_row, obs_rows), radius, epsilon)
    lw = np.sqrt(lweights[luse])                                  # wrapper.py:91
    return ketkf_weights(normed_perts[..., luse] * lw, normed_obs[..., luse] * lw, inf_factor, kernel)


# ----------------------------------------------------------------------------------------------
# IEnKS weight update  (pytassim/core/ienks.py, pytassim/core/utils.py:96-199)
# ----------------------------------------------------------------------------------------------


def svd(tensor: np.ndarray, reg_value: float = 0.0):
    """core/utils.py:96-131: ``torch.svd`` (tensor = u diag(s) v^T) with ``s + reg_value``."""
    u, s, vh = np.linalg.svd(tensor)
    return u, s + reg_value, vh.T


def rev_svd(u: np.ndarray, s: np.ndarray, v: np.ndarray) -> np.ndarray:
    """core/utils.py:134-150: ``(u * s) v^T``."""
    return (u * s) @ v.T


def ienks_update_weights(weights: np.ndarray, normed_perts: np.ndarray, normed_obs: np.ndarray, tau: float = 1.0,
                         epsilon: Optional[float] = None):
    """``IEnKSTransformModule._update_weights`` (core/ienks.py:117-132); ``epsilon`` not None: ``IEnKSBundleModule``
    (``_get_dh_dw`` :168-174).  weights (k, k), normed_perts (k, p), normed_obs (1, p) -> (w_mean (k, 1), w_perts (k, k))."""
    ens_size = weights.shape[-2]                                                      # :123
    weights_deviation = weights - np.eye(ens_size)                                    # :50 diagonal_add(weights, -1)
    w_mean = weights_deviation.mean(axis=1, keepdims=True)                            # :51
    w_perts = weights - w_mean                                                        # :52
    u, s, v = svd(w_perts)                                                            # :62
    s_inv = 1 / s                                                                     # :63
    s_prec = np.square(s_inv)                                                         # :64
    w_perts_inv = rev_svd(u, s_inv, v).T                                              # :65
    w_prec = rev_svd(u, s_prec, u) * (ens_size - 1)                                   # :66
    if epsilon is None:
        dh_dw = w_perts_inv @ normed_perts                                            # :74
    else:
        dh_dw = normed_perts / epsilon                                                # :173
    dlobs_dh = -normed_obs                                                            # :84
    grad_obs = dh_dw @ dlobs_dh.T                                                     # :85 matrix_product
    grad = (ens_size - 1) * w_mean + grad_obs                                         # :86-87
    new_prec = dh_dw @ dh_dw.T                                                        # :96
    new_prec = new_prec + (ens_size - 1.) * np.eye(ens_size)                          # :97
    updated_prec = (1 - tau) * w_prec + tau * new_prec                                # :98
    u, s, v = svd(updated_prec, reg_value=0.0)                                        # :99
    s_inv = 1 / s                                                                     # :100
    weights_cov = rev_svd(u, s_inv, v)                                                # :101
    s_perts = np.sqrt(s_inv * (ens_size - 1))                                         # :102
    weights_perts = rev_svd(u, s_perts, v)                                            # :103
    delta_weight = weights_cov @ grad                                                 # :130
    w_mean = w_mean - tau * delta_weight                                              # :131
    return w_mean, weights_perts


def ienks_weights(weights: np.ndarray, normed_perts: np.ndarray, normed_obs: np.ndarray, tau: float = 1.0,
                  epsilon: Optional[float] = None) -> np.ndarray:
    """``IEnKSTransformModule.forward`` / ``IEnKSBundleModule.forward`` (core/ienks.py:134-151): without observations the
    incoming weights are returned unchanged."""
    weights = np.asarray(weights, dtype=float)
    normed_perts = np.asarray(normed_perts, dtype=float)
    normed_obs = np.asarray(normed_obs, dtype=float)
    if normed_perts.shape[-1] != normed_obs.shape[-1]:                                # core/base.py:28-38
        raise ValueError('Observational size between ensemble ({0:d}) and observations '
                         '({1:d}) do not match!'.format(normed_perts.shape[-1], normed_obs.shape[-1]))
    weights = weights.reshape(-1, weights.shape[-1])                                  # :142
    if normed_perts.shape[-1] > 0:                                                    # :143
        w_mean, w_perts = ienks_update_weights(weights, normed_perts.reshape(-1, normed_perts.shape[-1]),
                                               normed_obs.reshape(1, -1), tau, epsilon)
        weights = w_mean + w_perts                                                    # :149
    return weights


def lienks_weights_point(grid_row, weights, normed_perts, normed_obs, obs_rows, dist_func, radius, tau=1.0, epsilon=None,
                         loc_epsilon=1e-5):
    """One grid point of the localized IEnKS (interface/lienks.py:68-118): ``localized_module`` with ``args_to_skip=(0,)``
    localizes the perturbations and innovations but hands the weights through (interface/wrapper.py:86-98)."""
    luse, lweights = gaspari_cohn_localize(dist_func(grid_row, obs_rows), radius, loc_epsilon)
    lw = np.sqrt(lweights[luse])
    return ienks_weights(weights, normed_perts[..., luse] * lw, normed_obs[..., luse] * lw, tau, epsilon)


# ----------------------------------------------------------------------------------------------
# Per-grid-point glue  (pytassim/interface/wrapper.py)
# ----------------------------------------------------------------------------------------------


def localized_weights(luse: np.ndarray, lweights: np.ndarray, normed_perts: np.ndarray,
                      normed_obs: np.ndarray, inf_factor: float) -> np.ndarray:
    """``wrapper_localization.localized_module`` after ``localize_obs`` returned (luse, lweights).

    Reference: interface/wrapper.py:86-98 — ``sqrt(w[use])`` scales the gathered columns of every
    argument; then the bridged module (wrapper.py:54-62) runs the core in float64."""
    lw = np.sqrt(lweights[luse])                                  # wrapper.py:91
    loc_perts = normed_perts[..., luse] * lw                      # :94-97
    loc_obs = normed_obs[..., luse] * lw
    return etkf_weights(loc_perts, loc_obs, inf_factor)


# ----------------------------------------------------------------------------------------------
# xarray glue restated on plain arrays  (pytassim/interface/base.py, observation.py, state.py)
# ----------------------------------------------------------------------------------------------


def split_mean_perts(array: np.ndarray, axis: int):
    """Reference: state.py:160-161."""
    mean = array.mean(axis=axis, keepdims=True)
    return mean, array - mean


def rcinv_diag(variance: np.ndarray) -> np.ndarray:
    """Reference: observation.py:241-245 — 1 / sqrt(var)."""
    return 1 / np.sqrt(variance)


def rcinv_chol(covariance: np.ndarray) -> np.ndarray:
    """Reference: observation.py:247-252 — inv(chol(R)^T)."""
    chol = np.linalg.cholesky(covariance).T
    return np.linalg.inv(chol)


def mul_rcinv(value: np.ndarray, covariance: np.ndarray) -> np.ndarray:
    """``Observation.mul_rcinv``: value (..., n_obs_grid) times R^{-1/2}.

    Reference: observation.py:273-295.  1-D covariance = variances -> elementwise ``value / sqrt(var)``;
    2-D -> ``xr.dot(value, inv(chol(R)^T), dims='obs_grid_1')`` i.e. ``value @ cinv``."""
    covariance = np.asarray(covariance, dtype=float)
    if covariance.ndim == 1:
        return value * rcinv_diag(covariance)                     # :277-279
    return value @ rcinv_chol(covariance)                         # :273-275


def obs_space_variables(ens_obs: Sequence[np.ndarray], observations: Sequence[np.ndarray],
                        covariances: Sequence[np.ndarray]):
    """``BaseAssimilation._get_obs_space_variables`` + ``_stack_obs`` on plain arrays.

    Reference: interface/base.py:359-379 and :223-241.
    ens_obs[i]       (k, n_t, n_i)   H(x) of obs dataset i (dims ensemble, time, obs_grid_1)
    observations[i]  (n_t, n_i)
    covariances[i]   (n_i,) variances or (n_i, n_i) covariance
    Returns innovations (M,), normalised perturbations (k, M) with M = sum_i n_t * n_i, stacked
    dataset-major, then time-major, then obs_grid_1 (the ``stack(obs_id=('time','obs_grid_1'))`` order)."""
    innovations: List[np.ndarray] = []
    perts: List[np.ndarray] = []
    for hx, y, cov in zip(ens_obs, observations, covariances):
        hx = np.asarray(hx, dtype=float)
        mean, pert = split_mean_perts(hx, axis=0)                 # base.py:367-369
        innov = np.asarray(y, dtype=float) - mean[0]              # :370
        innov = mul_rcinv(innov, cov)                             # :371
        pert = mul_rcinv(pert, cov)                               # :372
        innovations.append(innov.reshape(-1))                     # _stack_obs :231-233
        perts.append(pert.reshape(pert.shape[0], -1))
    return np.concatenate(innovations, axis=0), np.concatenate(perts, axis=1)   # :236


def apply_weights(state: np.ndarray, weights: np.ndarray) -> np.ndarray:
    """``BaseAssimilation._apply_weights``.  Reference: interface/base.py:257-278.

    state (n_var, n_t, k, N); weights (k, k) global or (N, k, k) per grid point
    (dims grid, ensemble, ensemble_new).  analysis[v,t,j,g] = mean[v,t,g] + sum_i perts[v,t,i,g] W[g,i,j]."""
    mean, perts = split_mean_perts(state, axis=2)                 # :267
    if weights.ndim == 2:
        ana_perts = np.einsum('vtig,ij->vtjg', perts, weights)    # :268 xr.dot over 'ensemble'
    else:
        ana_perts = np.einsum('vtig,gij->vtjg', perts, weights)
    return mean + ana_perts                                       # :270


# ----------------------------------------------------------------------------------------------
# Whole-path drivers
# ----------------------------------------------------------------------------------------------


def _localize(taper: str, dist, radius, epsilon):
    if taper == "gc":
        return gaspari_cohn_localize(dist, radius, epsilon)
    if taper == "gcinf":
        return gaspari_cohn_inf_localize(dist, float(np.atleast_1d(radius)[0]), epsilon)
    raise ValueError(taper)


def letkf_weights_point(grid_row, normed_perts, normed_obs, obs_rows, dist_func, radius,
                        epsilon=1e-5, inf_factor=1.0, taper="gc"):
    """One iteration of the reference's hot loop (interface/letkf.py:127-143 -> wrapper.py:86-98).
    Returns (W (k,k), indices of the local observations = np.nonzero(use)[0], weights[use])."""
    if dist_func is None:                                         # wrapper.py:87 localization None
        W = etkf_weights(normed_perts, normed_obs, inf_factor)
        idx = np.arange(normed_perts.shape[-1])
        return W, idx, np.ones(idx.shape)
    dist = dist_func(grid_row, obs_rows)
    luse, lweights = _localize(taper, dist, radius, epsilon)
    W = localized_weights(luse, lweights, normed_perts, normed_obs, inf_factor)
    return W, np.nonzero(luse)[0], lweights[luse]


def letkf_analysis(state, normed_perts, normed_obs, grid_rows, obs_rows, dist_func, radius,
                   epsilon=1e-5, inf_factor=1.0, taper="gc", grid_subset: Optional[Sequence[int]] = None,
                   return_lists: bool = False):
    """LETKF analysis = loop of ``letkf_weights_point`` over grid points + ``apply_weights``.

    state (n_var, n_t, k, N); normed_perts (k, M); normed_obs (M,); grid_rows (N, 1+nc); obs_rows (M, 1+nc).
    With ``grid_subset`` only those grid points are analysed (bounded CPU baseline samples) and the
    returned analysis has that many grid columns."""
    n_grid = state.shape[-1]
    sel = np.arange(n_grid) if grid_subset is None else np.asarray(grid_subset)
    k = state.shape[2]
    weights = np.empty((len(sel), k, k))
    lists = []
    for n, g in enumerate(sel):
        W, idx, _ = letkf_weights_point(grid_rows[g], normed_perts, normed_obs, obs_rows, dist_func,
                                        radius, epsilon, inf_factor, taper)
        weights[n] = W
        if return_lists:
            lists.append(idx)
    analysis = apply_weights(state[..., sel], weights)
    if return_lists:
        return analysis, weights, lists
    return analysis, weights


def etkf_analysis(state, normed_perts, normed_obs, inf_factor=1.0):
    """Global ETKF: one weight matrix for the whole state (interface/etkf.py:99-120 + base.py:257-278)."""
    W = etkf_weights(normed_perts, normed_obs, inf_factor)
    return apply_weights(state, W), W
