"""``BaseAssimilation`` / ``FilterAssimilation`` of the B200 engine: the reference's ``assimilate`` driver
(pytassim/interface/base.py:419-512) and ``update_state`` (pytassim/interface/filter.py:96-165) with the same
signatures, validation and error behaviour; everything between "obs-space variables are ready" and "analysis is
assembled" runs on the GPU (include/b200da.h).

Inputs are duck-typed (see :mod:`pytassim_b200.xrlite`): real ``xarray`` objects when xarray is installed, the
light-weight stand-ins otherwise.
"""
import abc
import logging
import time
import warnings

import numpy as np
import pandas as pd
import torch

logger = logging.getLogger(__name__)


class StateError(Exception):            # pytassim/state.py:44-49
    pass


class ObservationError(Exception):      # pytassim/observation.py:42-47
    pass


def _is_state_like(obj):
    return all(hasattr(obj, a) for a in ("values", "dims", "indexes"))


def _is_obs_like(obj):
    return hasattr(obj, "obs") and hasattr(obj, "__getitem__")


def dtindex_to_total_seconds(index):
    """pytassim/utilities/pandas.py:28-45."""
    index = pd.DatetimeIndex(index)
    return np.asarray((index - pd.Timestamp(1970, 1, 1)).total_seconds(), dtype=np.float64)


def to_host(tensor):
    """Device tensor -> numpy array in page-locked host memory: one DMA at PCIe speed instead of the driver's staged copy into
    pageable memory (400 MB of analysis: 180 -> ~10 ms).  The array keeps the pinned block alive; torch's pinned allocator
    reuses it once the caller drops the analysis."""
    if not tensor.is_cuda:
        return tensor.numpy()
    host = torch.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True)
    host.copy_(tensor, non_blocking=True)
    torch.cuda.current_stream(tensor.device).synchronize()
    return host.numpy()


def index_to_array(index):
    """pytassim/utilities/pandas.py:70-102: any index (incl. MultiIndex of tuples) -> float (n, n_coord) array."""
    if isinstance(index, pd.MultiIndex):
        return np.stack([np.asarray(index.get_level_values(i), dtype=np.float64) for i in range(index.nlevels)], axis=1)
    raw = np.atleast_1d(np.asarray(getattr(index, "values", index)))
    if raw.dtype == object and len(raw) and isinstance(raw[0], tuple):
        return np.asarray([list(t) for t in raw], dtype=np.float64)
    if raw.ndim > 1:
        return raw.astype(np.float64)
    return raw.astype(np.float64).reshape(-1, 1)


def _time_coord(obj):
    t = obj.indexes['time']
    return pd.DatetimeIndex(t) if not isinstance(t, pd.DatetimeIndex) else t


class BaseAssimilation(object):
    """Reference: pytassim/interface/base.py:51-512."""

    def __init__(self, smoother=False, gpu=False, pre_transform=None, post_transform=None, weight_save_path=None,
                 forward_model=None):
        self._dtype = torch.float32
        self.smoother = smoother
        self.gpu = gpu                    # accepted for signature compatibility; the engine always runs on the GPU
        self.pre_transform = pre_transform
        self.post_transform = post_transform
        self.dtype = torch.float64
        self.weight_save_path = weight_save_path
        self.forward_model = forward_model

    # -- properties (base.py:83-126) ---------------------------------------------------------------------------
    @property
    def core_module(self):
        """base.py:83-85: the device-backed core module of this algorithm (:mod:`pytassim_b200.core`)."""
        return self._make_core_module()

    def _make_core_module(self):
        raise NotImplementedError("this assimilation has no core module")

    @property
    def module(self):
        """base.py:87-104: the core module bridged to numpy arrays (interface/wrapper.py:29-62)."""
        from .per_point import NumpyBridge
        return NumpyBridge(self.core_module, self.device, self.dtype)

    @property
    def dtype(self):
        return self._dtype

    @dtype.setter
    def dtype(self, new_type):
        if isinstance(new_type, torch.dtype):
            self._dtype = new_type
        else:
            raise TypeError('Given object is not a valid torch.dtype, instead it has as type: {0}'.format(type(new_type)))

    @property
    def device(self):
        return torch.device("cuda")

    @property
    def chunks(self):
        return None

    # -- validation (base.py:129-151) ------------------------------------------------------------------------------
    @staticmethod
    def _validate_state(state):
        if not _is_state_like(state):
            raise TypeError('*** Given state is not a valid ``xarray.DataArray`` ***\n{0}'.format(type(state)))
        if tuple(state.dims) != ('var_name', 'time', 'ensemble', 'grid'):          # state.py:114-115
            raise StateError('*** Given state is not a valid state ***\n{0:s}'.format(str(state.dims)))

    @staticmethod
    def _validate_observations(observations):
        for obs in observations:
            if not _is_obs_like(obs):
                raise TypeError('*** Given observation is not a valid ``xarray.Dataset`` ***\n{0}'.format(type(obs)))
            if not obs.obs.valid:
                raise ObservationError('*** Given observation is not a valid observation ***\n{0:s}'.format(str(obs)))

    @staticmethod
    def _get_analysis_time(state, analysis_time=None):
        """base.py:154-178: None -> last time; exact match; else nearest with a UserWarning."""
        times = _time_coord(state)
        if analysis_time is None:
            return pd.Timestamp(times[-1])
        analysis_time = pd.to_datetime(analysis_time)
        if analysis_time in times:
            return pd.Timestamp(analysis_time)
        nearest = times[np.argmin(np.abs((times - analysis_time).total_seconds()))]
        warnings.warn('Given analysis time {0:s} is not within state, used instead nearest neighbor {1:s}'.format(
            str(analysis_time), str(nearest)), category=UserWarning)
        return pd.Timestamp(nearest)

    @staticmethod
    def _apply_obs_operator(pseudo_state, observations):
        """base.py:181-220: datasets whose operator raises NotImplementedError are dropped."""
        obs_equivalent, filtered = [], []
        for obs in observations:
            try:
                obs_equivalent.append(obs.obs.operator(obs, pseudo_state))
                filtered.append(obs)
            except NotImplementedError:
                pass
        return obs_equivalent, filtered

    def propagate_model(self, state, iter_num=0):
        """base.py:327-340 with the filter's prior weights (identity, filter.py:139-141): the model state handed to
        ``forward_model`` is ``mean + perturbations`` exactly as ``_apply_weights`` forms it (base.py:257-278)."""
        values = np.asarray(state.values)
        mean = values.mean(axis=2, keepdims=True)                       # state.py:160-161
        model_state = state.copy(data=mean + (values - mean))
        _, pseudo_state = self.forward_model(model_state, iter_num)
        self._validate_state(pseudo_state)
        return pseudo_state

    def get_pseudo_state(self, pseudo_state, state, iter_num=0):
        """base.py:342-357."""
        if pseudo_state is None and self.forward_model is not None:
            return self.propagate_model(state, iter_num)
        return state if pseudo_state is None else pseudo_state

    # -- obs-space variables (base.py:359-379, 223-241; observation.py:241-295) ---------------------------------------
    @staticmethod
    def _get_obs_space_variables(ens_obs, observations):
        """Returns (innovations (M,), perturbations (k, M), obs_info (M, 1+nc) = [t_unix, coords...]) stacked
        dataset-major, time-major, obs_grid_1-minor."""
        innovations, perts, infos = [], [], []
        for hx, obs in zip(ens_obs, observations):
            hxv = hx.transpose('ensemble', 'time', 'obs_grid_1').values.astype(np.float64) \
                if tuple(hx.dims) != ('ensemble', 'time', 'obs_grid_1') else np.asarray(hx.values, dtype=np.float64)
            y = np.asarray(obs['observations'].values, dtype=np.float64)
            cov = obs['covariance']
            covv = np.asarray(cov.values, dtype=np.float64)
            mean = hxv.mean(axis=0)                                     # state.py:160-161
            pert = hxv - mean[None]
            innov = y - mean
            if 'obs_grid_2' in cov.dims:                                # observation.py:247-275
                if 'time' in cov.dims:
                    cinv = np.stack([np.linalg.inv(np.linalg.cholesky(c).T) for c in covv], axis=0)
                    innov = np.einsum('to,top->tp', innov, cinv)
                    pert = np.einsum('kto,top->ktp', pert, cinv)
                else:
                    cinv = np.linalg.inv(np.linalg.cholesky(covv).T)
                    innov = innov @ cinv
                    pert = pert @ cinv
            else:                                                       # observation.py:241-245,277-279
                rc = 1 / np.sqrt(covv)
                innov = innov * rc
                pert = pert * rc
            t_unix = dtindex_to_total_seconds(_time_coord(obs['observations']))
            coords = index_to_array(obs['observations'].indexes['obs_grid_1'])
            n_t, n_o = y.shape
            info = np.concatenate([np.repeat(t_unix, n_o)[:, None], np.tile(coords, (n_t, 1))], axis=1)
            innovations.append(innov.reshape(-1))
            perts.append(pert.reshape(pert.shape[0], -1))
            infos.append(info)
        return np.concatenate(innovations), np.concatenate(perts, axis=1), np.concatenate(infos, axis=0)

    @staticmethod
    def _stack_obs_space_inputs(ens_obs, observations):
        """The raw inputs of ``_get_obs_space_variables`` stacked like ``_stack_obs`` (base.py:223-241) for the device
        prep kernel: (hx (k, M), y (M,), variance (M,), obs_info (M, 1+nc)), or None when a dataset carries a correlated
        covariance (obs_grid_2): that case is normalised on the host (observation.py:247-275)."""
        hxs, ys, vars_, infos = [], [], [], []
        for hx, obs in zip(ens_obs, observations):
            cov = obs['covariance']
            if 'obs_grid_2' in cov.dims:
                return None
            hxv = hx.transpose('ensemble', 'time', 'obs_grid_1').values \
                if tuple(hx.dims) != ('ensemble', 'time', 'obs_grid_1') else hx.values
            hxv = np.asarray(hxv, dtype=np.float64)
            y = np.asarray(obs['observations'].values, dtype=np.float64)
            n_t, n_o = y.shape
            covv = np.broadcast_to(np.asarray(cov.values, dtype=np.float64), (n_t, n_o))
            t_unix = dtindex_to_total_seconds(_time_coord(obs['observations']))
            coords = index_to_array(obs['observations'].indexes['obs_grid_1'])
            infos.append(np.concatenate([np.repeat(t_unix, n_o)[:, None], np.tile(coords, (n_t, 1))], axis=1))
            hxs.append(hxv.reshape(hxv.shape[0], -1)); ys.append(y.reshape(-1)); vars_.append(covv.reshape(-1))
        return np.concatenate(hxs, axis=1), np.concatenate(ys), np.concatenate(vars_), np.concatenate(infos, axis=0)

    @staticmethod
    def _stack_gather_inputs(pseudo_state, observations):
        """The inputs of ``b200da_obs_gather_prep`` when EVERY dataset's operator selects grid columns (it has
        ``device_index``, see :mod:`pytassim_b200.obs_ops`) and carries a diagonal R: (src_offset (M,) int64 into the
        flattened (n_var, n_time, k, N) pseudo state, member stride, y (M,), variance (M,), obs_info (M, 1+nc)) stacked like
        ``_stack_obs`` (base.py:223-241); None otherwise (the operators are then called on the host, base.py:181-220)."""
        if not observations or tuple(pseudo_state.dims) != ('var_name', 'time', 'ensemble', 'grid'):
            return None
        for obs in observations:
            if not callable(getattr(obs.obs.operator, 'device_index', None)) or 'obs_grid_2' in obs['covariance'].dims:
                return None
        _, n_time, k, n_grid = pseudo_state.values.shape
        srcs, ys, vars_, infos = [], [], [], []
        for obs in observations:
            var_pos, t_pos, g_pos = obs.obs.operator.device_index(obs, pseudo_state)
            y = np.asarray(obs['observations'].values, dtype=np.float64)
            n_t, n_o = y.shape
            covv = np.broadcast_to(np.asarray(obs['covariance'].values, dtype=np.float64), (n_t, n_o))
            base = (var_pos * n_time + t_pos) * (k * n_grid)                                # member 0 of (variable, time)
            srcs.append((base[:, None] + g_pos[None, :]).reshape(-1))
            t_unix = dtindex_to_total_seconds(_time_coord(obs['observations']))
            coords = index_to_array(obs['observations'].indexes['obs_grid_1'])
            infos.append(np.concatenate([np.repeat(t_unix, n_o)[:, None], np.tile(coords, (n_t, 1))], axis=1))
            ys.append(y.reshape(-1)); vars_.append(covv.reshape(-1))
        return (np.concatenate(srcs).astype(np.int64), int(n_grid), np.concatenate(ys), np.concatenate(vars_),
                np.concatenate(infos, axis=0))

    # -- weight store (base.py:280-324) ------------------------------------------------------------------------------
    def store_weights(self, weights):
        """base.py:280-300: ``weights`` (DataArray-like with dims ['grid',] 'ensemble', 'ensemble_new') -> netCDF under
        ``weight_save_path``; MultiIndex dimensions are stored as single-dimensional indexes (utilities/xarray.py:58-90)."""
        from ..utilities import save_netcdf
        return save_netcdf(dataset_to_save=weights, save_path=self.weight_save_path)

    def load_weights(self):
        """base.py:302-324: the stored weights with their MultiIndexes decoded (the reference's dask chunking of the loaded
        array has no counterpart here: the weights go back to the device in one piece)."""
        from ..utilities import load_netcdf
        return load_netcdf(load_path=self.weight_save_path, array=True)

    def _weights_through_store(self, state, weights):
        """filter.py:159-162: store the estimated weights, load them back, hand the loaded values on to the update.
        ``weights``: host array (k, k) (global) or (N, k, k) (one matrix per grid point, reference dims grid x ensemble x
        ensemble_new, interface/letkf.py:143-147)."""
        from ..xrlite import DataArray
        ens = state.indexes['ensemble']
        coords = dict(ensemble=ens, ensemble_new=ens)
        dims = ('ensemble', 'ensemble_new')
        if weights.ndim == 3:
            coords['grid'] = state.indexes['grid']
            dims = ('grid', ) + dims
        self.store_weights(DataArray(np.asarray(weights, dtype=np.float64), coords, dims))
        loaded = self.load_weights()
        if loaded.dims != dims:
            loaded = loaded.transpose(*dims)
        return np.ascontiguousarray(loaded.values)

    @abc.abstractmethod
    def update_state(self, state, observations, pseudo_state, analysis_time):
        pass

    def assimilate(self, state, observations, pseudo_state=None, analysis_time=None):
        """Reference: pytassim/interface/base.py:419-512 (same signature, warnings and exceptions)."""
        start_time = time.time()
        logger.info('Starting assimilation')
        if not observations:
            warnings.warn('No observation is given, I will return the background state!', UserWarning)
            return state
        if not isinstance(observations, (list, set, tuple)):
            observations = (observations, )
        self._validate_state(state)
        self._validate_observations(observations)
        analysis_time = self._get_analysis_time(state, analysis_time)
        if self.pre_transform:
            for trans in self.pre_transform:
                state, observations, pseudo_state = trans.pre(state, observations, pseudo_state)
        analysis = self.update_state(state, observations, pseudo_state, analysis_time)
        if self.post_transform:
            for trans in self.post_transform:
                analysis = trans.post(analysis, state, observations, pseudo_state)
        self._validate_state(analysis)
        logger.info('Finished assimilation after {0:.2f} s'.format(time.time() - start_time))
        return analysis


class FilterAssimilation(BaseAssimilation):
    """Reference: pytassim/interface/filter.py:29-165."""

    @staticmethod
    def _slice_analysis(analysis_time, state, observations, pseudo_state):
        """filter.py:39-54: state / pseudo state / observations at the analysis time only."""
        def pick(obj):
            times = _time_coord(obj)
            hit = np.nonzero(times == analysis_time)[0]
            if len(hit) == 0:
                raise KeyError(analysis_time)
            return hit[:1]
        state = state.isel(time=pick(state))
        pseudo_state = pseudo_state.isel(time=pick(pseudo_state))
        sel_obs = []
        for obs in observations:
            tmp = obs.isel(time=pick(obs['observations']))
            tmp.obs.operator = obs.obs.operator
            sel_obs.append(tmp)
        return state, sel_obs, pseudo_state

    @abc.abstractmethod
    def _analyse_arrays(self, state, x, innov, perts, obs_info):
        """x (n_slices, k, N) host array, innov (M,) / perts (k, M) host arrays or device tensors ->
        analysed (n_slices, k, N) host array."""

    @abc.abstractmethod
    def _prep_engine(self, k, n_slices):
        """An engine of the right ensemble size / dtype whose ``obs_prep`` runs the device prep kernel."""

    def update_state(self, state, observations, pseudo_state, analysis_time):
        pseudo_state = self.get_pseudo_state(pseudo_state, state)
        self._validate_state(pseudo_state)
        if not self.smoother:
            state, observations, pseudo_state = self._slice_analysis(analysis_time, state, observations, pseudo_state)
        values = np.ascontiguousarray(state.values, dtype=np.float64)
        n_var, n_t, k, n_grid = values.shape
        gathered = self._stack_gather_inputs(pseudo_state, observations)
        if gathered is not None:                    # column-selecting operators + diagonal R: operator and prep on the device
            src, member_stride, y, var, obs_info = gathered
            pv = pseudo_state.values
            same = pv.shape == values.shape and (pseudo_state is state or np.shares_memory(pv, values, max_work=1)
                                                 or np.array_equal(pv, values))
            values = torch.as_tensor(values).cuda()  # one upload serves the operator gather and the analysis
            xp = values if same else np.ascontiguousarray(pseudo_state.values, dtype=np.float64)
            perts, innov = self._prep_engine(k, n_var * n_t).obs_gather_prep(xp, src, member_stride, y, var)
        else:
            ens_obs, filtered_obs = self._apply_obs_operator(pseudo_state, observations)
            if not filtered_obs:
                warnings.warn('No observation is given, I will return the background state!', UserWarning)
                return state
            stacked = self._stack_obs_space_inputs(ens_obs, filtered_obs)
            if stacked is not None:                 # diagonal R: mean / perturbations / innovations / R^-1/2 on the device
                hx, y, var, obs_info = stacked
                perts, innov = self._prep_engine(k, n_var * n_t).obs_prep(hx, y, var)
            else:
                innov, perts, obs_info = self._get_obs_space_variables(ens_obs, filtered_obs)
        xa = self._analyse_arrays(state, values.reshape(n_var * n_t, k, n_grid), innov, perts, obs_info)
        return state.copy(data=np.asarray(xa).reshape(state.values.shape).astype(state.values.dtype, copy=False))
