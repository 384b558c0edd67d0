"""``wrapper_bridge`` / ``wrapper_localization`` of the reference (pytassim/interface/wrapper.py:29-99) around the device-backed
core modules: numpy in, numpy out, one grid point per call.  The ``assimilate`` path does not use them (all grid points are
analysed by one launch); they exist for code that calls ``assimilation.module`` / ``.localized_module`` directly."""
import numpy as np
import torch

__all__ = ['wrapper_bridge', 'wrapper_localization']


def wrapper_bridge(core_module, device, dtype):
    def bridged_module(*args):                                                      # wrapper.py:54-62
        torch_args = [torch.from_numpy(np.ascontiguousarray(arg)).to(device=device, dtype=dtype) for arg in args]
        torch_weights = core_module(*torch_args).cpu().detach()
        return torch_weights.numpy().astype(args[0].dtype)
    return bridged_module


def wrapper_localization(module, localization):
    def localized_module(grid_info, *args, obs_info=None, args_to_skip=None):       # wrapper.py:86-98
        if localization is not None:
            luse, lweights = localization.localize_obs(grid_info, obs_info)
            lweights = np.sqrt(lweights[luse])
            if args_to_skip is None:
                args_to_skip = []
            args = [arg if k in args_to_skip else arg[..., luse] * lweights for k, arg in enumerate(args)]
        return module(*args)
    return localized_module
