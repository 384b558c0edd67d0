"""Global ETKF on the B200 engine.  Reference: pytassim/interface/etkf.py:43-120."""
import torch

from .base import FilterAssimilation, to_host
from ..engine import LETKFEngine
from ..localization.metrics import AbsDistance1D

__all__ = ['ETKF']


class ETKF(FilterAssimilation):
    def __init__(self, inf_factor=1.0, smoother=False, gpu=False, pre_transform=None, post_transform=None,
                 weight_save_path=None, forward_model=None):
        super().__init__(smoother=smoother, gpu=gpu, pre_transform=pre_transform, post_transform=post_transform,
                         weight_save_path=weight_save_path, forward_model=forward_model)
        self.inf_factor = inf_factor
        self._engines = {}

    def __str__(self):
        return 'Global ETKF(inf_factor={0})'.format(str(self.inf_factor.item()))

    def __repr__(self):
        return 'ETKF({0})'.format(repr(self.inf_factor.item()))

    @property
    def inf_factor(self):
        return self._inf_factor

    @inf_factor.setter
    def inf_factor(self, new_factor):                      # etkf.py:93-97
        if isinstance(new_factor, (float, int)):
            new_factor = torch.tensor(new_factor, dtype=self.dtype)
        self._inf_factor = new_factor
        self._engines = {}

    def _make_core_module(self):
        from ..core import ETKFModule
        return ETKFModule(inf_factor=self.inf_factor)

    def _kernel_key(self):
        """Part of the engine cache key that identifies the ensemble-space kernel (KETKF / LKETKF override it)."""
        return ()

    def _configure_engine(self, engine):
        """Hook for subclasses: called once on every freshly created engine (KETKF / LKETKF set the kernel program)."""
        return engine

    def _global_engine(self, k, n_slices):
        key = ('global', k, n_slices, float(self.inf_factor), self.dtype, self._kernel_key())
        if key not in self._engines:
            self._engines[key] = self._configure_engine(
                LETKFEngine(k, n_slices, AbsDistance1D(), 1.0, inf_factor=float(self.inf_factor), dtype=self.dtype))
        return self._engines[key]

    def _prep_engine(self, k, n_slices):
        return self._global_engine(k, n_slices)

    def _analyse_arrays(self, state, x, innov, perts, obs_info):
        eng = self._global_engine(x.shape[1], x.shape[0])
        weights = eng.etkf_weights(perts, innov)                            # etkf.py:99-120
        if self.weight_save_path is not None:                               # filter.py:159-162
            weights = self._weights_through_store(state, weights.cpu().numpy())
        return to_host(eng.apply_weights(torch.as_tensor(x), weights))  # base.py:257-278
