"""Core modules of the reference (pytassim/core/*.py) as device-backed callables.

The reference's plug-in point (3) is ``core_module(normed_perts, normed_obs) -> Tensor(k, k)``: a
``torch.nn.Module`` that computes one weight matrix.  The classes here keep the names, constructor arguments and call
signatures of ``ETKFModule`` (core/etkf.py:28-103), ``KETKFModule`` (core/ketkf.py:28-100), ``IEnKSTransformModule`` and
``IEnKSBundleModule`` (core/ienks.py:28-174); a call runs the Gram, (kernelise / IEnKS pre-pass) and ensemble-space solve
kernels of ``libb200da.so`` for that one problem and returns a tensor on the device.  They are what ``assimilation.core_module``
/ ``.module`` / ``.localized_module`` hand out; the ``assimilate`` path itself never goes through them (it analyses all grid
points in one launch).  There is no CPU path.
"""
import torch

from ..engine import LETKFEngine
from ..localization.metrics import AbsDistance1D

__all__ = ['BaseModule', 'ETKFModule', 'KETKFModule', 'IEnKSTransformModule', 'IEnKSBundleModule']


class BaseModule(object):
    """core/base.py:26-62 (size check, 2-d views); engines are cached per (ensemble size, dtype)."""

    def __init__(self):
        self._engines = {}

    @staticmethod
    def _test_sizes(normed_perts, normed_obs):
        if normed_perts.shape[-1] != normed_obs.shape[-1]:                          # core/base.py:33-38
            raise ValueError('Observational size between ensemble ({0:d}) and observations '
                             '({1:d}) do not match!'.format(normed_perts.shape[-1], normed_obs.shape[-1]))

    @staticmethod
    def _view_as_2d(tensor):
        tensor = torch.as_tensor(tensor)
        return tensor.reshape(1, -1) if tensor.dim() < 2 else tensor.reshape(-1, tensor.shape[-1])   # core/base.py:41-46

    def _configure(self, engine):
        return engine

    def _inflation(self):
        return 1.0

    def _config_key(self):
        """What ``_configure`` puts on the engine (the kernel program of KETKFModule): part of the cache key, so reassigning
        ``module.kernel`` or changing a kernel's parameters takes effect on the next call, as in the reference."""
        return None

    def _engine(self, ens_size, dtype):
        key = (int(ens_size), dtype, float(self._inflation()), self._config_key())
        if key not in self._engines:
            self._engines = {key: self._configure(LETKFEngine(int(ens_size), 1, AbsDistance1D(), 1.0,
                                                              inf_factor=float(self._inflation()), dtype=dtype))}
        return self._engines[key]

    def __call__(self, *args):
        return self.forward(*args)

    def to(self, *args, **kwargs):
        return self


class ETKFModule(BaseModule):
    """core/etkf.py:28-103."""

    def __init__(self, inf_factor=1.0):
        super().__init__()
        self.inf_factor = inf_factor

    def __str__(self):
        return 'ETKFCore({0})'.format(self.inf_factor)

    def __repr__(self):
        return 'ETKFCore'

    def _inflation(self):
        return float(self.inf_factor)

    def forward(self, normed_perts, normed_obs):
        normed_perts, normed_obs = torch.as_tensor(normed_perts), torch.as_tensor(normed_obs)
        self._test_sizes(normed_perts, normed_obs)
        dtype = normed_perts.dtype if normed_perts.dtype in (torch.float32, torch.float64) else torch.float64
        perts = self._view_as_2d(normed_perts) if normed_perts.shape[-1] > 0 else normed_perts.reshape(normed_perts.shape[-2], 0)
        return self._engine(perts.shape[0], dtype).etkf_weights(perts, normed_obs.reshape(-1))


class KETKFModule(ETKFModule):
    """core/ketkf.py:28-100: ``kernel`` is a :mod:`pytassim_b200.kernels` descriptor."""

    def __init__(self, kernel, inf_factor=1.0):
        super().__init__(inf_factor)
        self.kernel = kernel

    def __str__(self):
        return 'KETKFModule({0:s}, {1})'.format(str(self.kernel), self.inf_factor)

    def __repr__(self):
        return 'KETKF({0:s})'.format(repr(self.kernel))

    def _configure(self, engine):
        return engine.set_kernel(self.kernel)

    def _config_key(self):
        k = self.kernel
        if k is None or getattr(k, "is_linear", False) or not hasattr(k, "program"):
            return None
        return tuple((int(op), float(a), float(b)) for op, a, b in k.program())


class IEnKSTransformModule(BaseModule):
    """core/ienks.py:28-151."""

    def __init__(self, tau=1.0):
        super().__init__()
        self.tau = tau

    def __str__(self):
        return 'TransformModule(tau={0})'.format(self.tau)

    def __repr__(self):
        return 'TransformModule'

    _epsilon = None

    def forward(self, weights, normed_perts, normed_obs):
        weights, normed_perts, normed_obs = (torch.as_tensor(a) for a in (weights, normed_perts, normed_obs))
        self._test_sizes(normed_perts, normed_obs)
        weights = self._view_as_2d(weights)
        perts = self._view_as_2d(normed_perts) if normed_perts.shape[-1] > 0 else normed_perts.reshape(weights.shape[-1], 0)
        dtype = weights.dtype if weights.dtype in (torch.float32, torch.float64) else torch.float64
        eps = None if self._epsilon is None else float(self._epsilon)
        return self._engine(weights.shape[-1], dtype).ienks_weights(weights, perts, normed_obs.reshape(-1), tau=float(self.tau),
                                                                  epsilon=eps)


class IEnKSBundleModule(IEnKSTransformModule):
    """core/ienks.py:154-174."""

    def __init__(self, epsilon=1E-4, tau=1.0):
        super().__init__(tau=tau)
        self.epsilon = epsilon

    def __str__(self):
        return 'IEnKSBundleModule(eps={0}, tau={1}'.format(str(self.epsilon), str(self.tau))

    def __repr__(self):
        return 'IEnKSBundle({0}, {1})'.format(repr(self.epsilon), repr(self.tau))

    @property
    def _epsilon(self):
        return self.epsilon
