"""Per-grid-point call forms of the device-backed core modules.

The reference exposes its hot loop as two closures (pytassim/interface/wrapper.py:29-99): ``assimilation.module`` takes numpy
arrays and returns the (k, k) weights of one problem, ``assimilation.localized_module`` additionally selects and tapers the
local observations of one grid row.  ``assimilate`` here never goes through them (all grid points are analysed by one
launch); these callables exist for code that drives the per-point interface directly, and every call runs on the device
(neighbour search for the selection, Gram + solve for the weights).
"""
import numpy as np
import torch

__all__ = ['NumpyBridge', 'LocalObservations']


class NumpyBridge(object):
    """numpy in, numpy out around a core module (what ``wrapper_bridge(core_module, device, dtype)`` returns in the
    reference): the result has the dtype of the first argument."""

    def __init__(self, core_module, device, dtype):
        self.core_module, self.device, self.dtype = core_module, device, dtype

    def __call__(self, *arrays):
        on_device = [torch.as_tensor(np.ascontiguousarray(a), dtype=self.dtype).to(self.device) for a in arrays]
        weights = self.core_module(*on_device)
        return weights.detach().cpu().numpy().astype(arrays[0].dtype)


class LocalObservations(object):
    """One grid point per call (what ``wrapper_localization(module, localization)`` returns in the reference):
    ``f(grid_info, *args, obs_info=..., args_to_skip=...)`` keeps the observations the localization selects for ``grid_info``,
    scales them with the square root of their weights and calls ``module``; arguments listed in ``args_to_skip`` (the incoming
    weights of the IEnKS, interface/lienks.py:109-112) pass through untouched.  Without a localization it is ``module``."""

    def __init__(self, module, localization):
        self.module, self.localization = module, localization

    def __call__(self, grid_info, *args, obs_info=None, args_to_skip=None):
        if self.localization is None:
            return self.module(*args)
        use, taper = self.localization.localize_obs(grid_info, obs_info)
        scale = np.sqrt(taper[use])
        skip = set(args_to_skip or ())
        local_args = [a if pos in skip else a[..., use] * scale for pos, a in enumerate(args)]
        return self.module(*local_args)
