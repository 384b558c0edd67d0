"""Iterative ensemble Kalman smoother (IEnKS) on the B200 engine: global and localized, transform and bundle variants.

Reference: pytassim/interface/variational.py:33-135 (``VarAssimilation``: the outer loop over ``max_iter`` iterations),
interface/ienks.py:34-164 (``IEnKSTransform``, ``IEnKSBundle``), interface/lienks.py:40-163 (``LocalizedIEnKSTransform``,
``LocalizedIEnKSBundle``), core/ienks.py:28-174 (the weight update).  Per iteration the host runs the user's ``forward_model``
and observation operators (as the reference does); the model state ``x_mean + X' W_model`` that is handed to the model, the
observation-space preparation, the weight update of every grid point (Gram -> k_ienks_pre -> ensemble-space solve) and the
final update run on the device.
"""
import abc

import numpy as np
import torch

from .base import BaseAssimilation, index_to_array
from ..engine import LETKFEngine
from ..localization.metrics import AbsDistance1D

__all__ = ['VarAssimilation', 'IEnKSTransform', 'IEnKSBundle', 'LocalizedIEnKSTransform', 'LocalizedIEnKSBundle']


def _bounded(value, dtype, min_val=None, max_val=None):
    """utilities/decorators.py:51-75 (``ensure_tensor`` + ``bound_tensor``)."""
    if not isinstance(value, torch.Tensor):
        value = torch.tensor(value, dtype=dtype)
    if (min_val is not None) and torch.any(value < min_val):
        raise ValueError('Given new value {0} is smaller than the minimum value {1}!'.format(value, min_val))
    if (max_val is not None) and torch.any(value > max_val):
        raise ValueError('Given new value {0} is larger than the maximum value {1}!'.format(value, max_val))
    return value


class VarAssimilation(BaseAssimilation):
    """interface/variational.py:33-135."""

    def __init__(self, forward_model, max_iter=10, smoother=False, gpu=False, pre_transform=None, post_transform=None,
                 weight_save_path=None):
        super().__init__(smoother=smoother, gpu=gpu, pre_transform=pre_transform, post_transform=post_transform,
                         forward_model=forward_model, weight_save_path=weight_save_path)
        self.max_iter = max_iter
        self._engines = {}

    # -- engines ---------------------------------------------------------------------------------------------------
    def _global_engine(self, k, n_slices):
        key = ('global', k, n_slices, self.dtype)
        if key not in self._engines:
            self._engines[key] = LETKFEngine(k, n_slices, AbsDistance1D(), 1.0, dtype=self.dtype)
        return self._engines[key]

    def _prep_engine(self, k, n_slices):
        return self._global_engine(k, n_slices)

    # -- pieces of the outer loop ----------------------------------------------------------------------------------------
    def _get_model_weights(self, weights):
        """base.py:326-327: the weights the model state is built with (the bundle variant overrides this)."""
        return weights

    def precompute_weights(self, state, weights):
        """variational.py:56-80: the reference stores the weights of every iteration and loads them again (under
        ``weight_save_path`` or a temporary file) to cut its dask graph; the values are unchanged.  Here the weights stay on
        the device unless ``weight_save_path`` is set, in which case they take the same netCDF round trip."""
        if self.weight_save_path is None:
            return weights
        loaded = self._weights_through_store(state, weights.cpu().numpy())
        return torch.as_tensor(loaded).to(weights.device)

    def _model_state(self, engine, state, x_dev, weights):
        """base.py:335-336: ``_apply_weights(state, model_weights)`` on the device -> DataArray-like for ``forward_model``."""
        model_weights = self._get_model_weights(weights)
        values = engine.apply_weights(x_dev, model_weights).cpu().numpy()
        return state.copy(data=values.reshape(state.values.shape).astype(state.values.dtype, copy=False))

    def _obs_space(self, engine, pseudo_state, observations):
        """base.py:181-220, 359-379: (perturbations (k, M), innovations (M,), obs_info (M, 1 + nc)); diagonal R on the device."""
        ens_obs, filtered_obs = self._apply_obs_operator(pseudo_state, observations)
        if not filtered_obs:
            return None
        stacked = self._stack_obs_space_inputs(ens_obs, filtered_obs)
        if stacked is not None:
            hx, y, var, obs_info = stacked
            perts, innov = engine.obs_prep(hx, y, var)
        else:
            innov, perts, obs_info = self._get_obs_space_variables(ens_obs, filtered_obs)
        return perts, innov, obs_info

    @abc.abstractmethod
    def inner_loop(self, engine, state, x_dev, weights, perts, innov, obs_info):
        """variational.py:82-90 on arrays: weights (k, k) or (N, k, k) device tensor -> updated weights."""

    def _select_analysis_time(self, state, analysis_time):
        times = np.asarray(state.indexes['time'])
        hit = np.nonzero(times == np.datetime64(analysis_time))[0] if np.issubdtype(times.dtype, np.datetime64) \
            else np.nonzero(times == analysis_time)[0]
        if len(hit) == 0:
            raise KeyError(analysis_time)
        return state.isel(time=hit[:1])                                              # variational.py:113

    def update_state(self, state, observations, pseudo_state, analysis_time):
        """variational.py:105-135."""
        state = self._select_analysis_time(state, analysis_time)
        values = np.ascontiguousarray(state.values, dtype=np.float64)
        n_var, n_t, k, n_grid = values.shape
        engine = self._analysis_engine(state, k, n_var * n_t)
        tdtype = self.dtype
        x_dev = torch.as_tensor(values.reshape(n_var * n_t, k, n_grid)).to(device=engine.device, dtype=tdtype)
        weights = torch.eye(k, dtype=tdtype, device=engine.device)                    # generate_prior_weights (base.py:243-255)
        iter_num = 0
        while iter_num < self.max_iter:
            if pseudo_state is None and self.forward_model is not None:              # base.py:342-357
                _, pseudo_state = self.forward_model(self._model_state(engine, state, x_dev, weights), iter_num)
                self._validate_state(pseudo_state)
            elif pseudo_state is None:
                pseudo_state = state
            obs_space = self._obs_space(engine, pseudo_state, observations)
            if obs_space is not None:
                perts, innov, obs_info = obs_space
                weights = self.inner_loop(engine, state, x_dev, weights, perts, innov, obs_info)
            weights = self.precompute_weights(state, weights)
            pseudo_state = None                                                      # variational.py:126
            iter_num += 1
        xa = engine.apply_weights(x_dev, weights).cpu().numpy()                       # variational.py:131
        analysis = state.copy(data=xa.reshape(state.values.shape).astype(state.values.dtype, copy=False))
        if self.smoother:
            analysis, _ = self.forward_model(analysis, iter_num)                     # variational.py:132-133
        return analysis

    def _analysis_engine(self, state, k, n_slices):
        return self._global_engine(k, n_slices)


class IEnKSTransform(VarAssimilation):
    """interface/ienks.py:34-118."""

    def __init__(self, forward_model, tau=1.0, max_iter=10, smoother=False, gpu=False, pre_transform=None, post_transform=None,
                 weight_save_path=None):
        super().__init__(forward_model=forward_model, max_iter=max_iter, smoother=smoother, gpu=gpu, pre_transform=pre_transform,
                         post_transform=post_transform, weight_save_path=weight_save_path)
        self.tau = tau

    def __str__(self):
        return 'IEnKSTransform(tau={0})'.format(str(self.tau.item()))

    def __repr__(self):
        return 'IEnKSTransform({0})'.format(repr(self.tau.item()))

    @property
    def tau(self):
        return self._tau

    @tau.setter
    def tau(self, new_tau):                                                          # ienks.py:88-94
        tau = _bounded(new_tau, self.dtype, 0.0, 1.0)
        if not float(tau) > 0.0:
            # the reference accepts tau = 0 (a step that leaves the mean weights where they are); the device kernels solve with
            # inflation 1 / tau (csrc/ienks_kernel.cuh), so reject it here, where it is set, instead of failing in assimilate()
            raise ValueError("tau = 0 (no update of the mean weights) is not supported by the B200 IEnKS kernels: "
                             "use a learning rate in (0, 1]")
        self._tau = tau

    @property
    def _epsilon_value(self):
        return None

    def _make_core_module(self):
        from ..core import IEnKSTransformModule
        return IEnKSTransformModule(tau=self.tau)

    def inner_loop(self, engine, state, x_dev, weights, perts, innov, obs_info):     # ienks.py:96-118
        return engine.ienks_weights(weights, perts, innov, tau=float(self.tau), epsilon=self._epsilon_value)


class IEnKSBundle(IEnKSTransform):
    """interface/ienks.py:121-164."""

    def __init__(self, forward_model, tau=1.0, epsilon=1E-4, max_iter=10, smoother=False, gpu=False, pre_transform=None,
                 post_transform=None, weight_save_path=None):
        super().__init__(forward_model=forward_model, tau=tau, max_iter=max_iter, smoother=smoother, gpu=gpu,
                         pre_transform=pre_transform, post_transform=post_transform, weight_save_path=weight_save_path)
        self.epsilon = epsilon

    def __str__(self):
        return 'IEnKSBundle(epsilon={0}, tau={1})'.format(str(self.epsilon.item()), str(self.tau.item()))

    def __repr__(self):
        return 'IEnKSBundle({0},{1})'.format(repr(self.epsilon.item()), repr(self.tau.item()))

    @property
    def epsilon(self):
        return self._epsilon

    @epsilon.setter
    def epsilon(self, new_epsilon):                                                  # ienks.py:131-138
        self._epsilon = _bounded(new_epsilon, self.dtype, 0.0, None)

    @property
    def _epsilon_value(self):
        return float(self.epsilon)

    def _make_core_module(self):
        from ..core import IEnKSBundleModule
        return IEnKSBundleModule(epsilon=self.epsilon, tau=self.tau)

    def _get_model_weights(self, weights):
        """ienks.py:153-160: ``epsilon * I + mean over ensemble_new`` — the bundle around the current mean weights."""
        k = weights.shape[-1]
        weights_mean = weights.mean(dim=-1, keepdim=True)                            # (k, 1) or (N, k, 1)
        return float(self.epsilon) * torch.eye(k, dtype=weights.dtype, device=weights.device) + weights_mean


class _LocalizedMixin(object):
    """interface/lienks.py:68-118 + interface/mixin_local.py:33-69 on arrays."""

    @property
    def chunks(self):
        return dict(grid=self.chunksize)

    @property
    def localized_module(self):
        """mixin_local.py:37-42."""
        from .per_point import LocalObservations
        return LocalObservations(module=self.module, localization=self.localization)

    def _analysis_engine(self, state, k, n_slices):
        loc = self.localization
        if loc is None:                                    # wrapper.py:87: without a localization every grid point sees every observation
            return self._global_engine(k, n_slices)
        nc = loc.dist_func.n_coord
        if getattr(loc.dist_func, 'zero_coords', False):
            grid_coords = np.zeros((len(state.indexes['grid']), nc))
        else:
            grid_coords = index_to_array(state.indexes['grid'])                      # mixin_local.py:50-69
            if grid_coords.shape[1] < nc:
                raise ValueError("the metric needs {0} coordinate column(s)".format(nc))
        grid_coords = np.ascontiguousarray(grid_coords[:, :nc])
        key = ('local', k, n_slices, self.dtype, type(loc).__name__, loc.dist_func.cache_key(), tuple(np.atleast_1d(loc.radius).tolist()),
               float(loc.epsilon))
        if key not in self._engines:
            self._engines = {kk: v for kk, v in self._engines.items() if kk[0] != 'local'}
            self._engines[key] = LETKFEngine(k, n_slices, loc.dist_func, loc.radius, epsilon=loc.epsilon, taper=loc.taper,
                                             dtype=self.dtype)
            self._grid_cache = None
        eng = self._engines[key]
        cache = getattr(self, '_grid_cache', None)
        if cache is None or cache.shape != grid_coords.shape or not np.array_equal(cache, grid_coords):
            eng.set_grid(grid_coords)
            self._grid_cache = grid_coords.copy()
        return eng

    def inner_loop(self, engine, state, x_dev, weights, perts, innov, obs_info):
        loc = self.localization
        if loc is None:
            return engine.ienks_weights(weights, perts, innov, tau=float(self.tau), epsilon=self._epsilon_value)
        nc = loc.dist_func.n_coord
        if getattr(loc.dist_func, 'zero_coords', False):
            obs_coords = np.zeros((obs_info.shape[0], nc))
        else:
            if obs_info.shape[1] - 1 < nc:
                raise ValueError("the metric needs {0} coordinate column(s)".format(nc))
            obs_coords = obs_info[:, 1:1 + nc]
        engine.bin_obs(obs_coords, perts, innov)
        _, new_weights = engine.ienks_step(x_dev, weights, tau=float(self.tau), epsilon=self._epsilon_value)
        return new_weights


class LocalizedIEnKSTransform(_LocalizedMixin, IEnKSTransform):
    """interface/lienks.py:40-118."""

    def __init__(self, forward_model, localization=None, tau=1.0, max_iter=10, smoother=False, gpu=False, pre_transform=None,
                 post_transform=None, chunksize=10, weight_save_path=None):
        super().__init__(forward_model=forward_model, tau=tau, max_iter=max_iter, smoother=smoother, gpu=gpu,
                         pre_transform=pre_transform, post_transform=post_transform, weight_save_path=weight_save_path)
        self.localization = localization
        self.chunksize = chunksize

    def __str__(self):
        return 'Localized IEnKSTransform(loc={0}, tau={1})'.format(str(self.localization), str(self.tau.item()))

    def __repr__(self):
        return 'LIEnKSTransform({0},{1})'.format(repr(self.localization), repr(self.tau.item()))


class LocalizedIEnKSBundle(_LocalizedMixin, IEnKSBundle):
    """interface/lienks.py:121-163."""

    def __init__(self, forward_model, localization=None, tau=1.0, epsilon=1E-4, max_iter=10, smoother=False, gpu=False,
                 pre_transform=None, post_transform=None, chunksize=10, weight_save_path=None):
        super().__init__(forward_model=forward_model, tau=tau, epsilon=epsilon, max_iter=max_iter, smoother=smoother, gpu=gpu,
                         pre_transform=pre_transform, post_transform=post_transform, weight_save_path=weight_save_path)
        self.localization = localization
        self.chunksize = chunksize

    def __str__(self):
        return 'Localized IEnKSBundle(loc={0}, eps={1}, tau={2})'.format(str(self.localization), str(self.epsilon.item()),
                                                                         str(self.tau.item()))

    def __repr__(self):
        return 'LIEnKSBundle({0},{1},{2})'.format(repr(self.localization), repr(self.epsilon.item()), repr(self.tau.item()))
