"""Kernelised ETKF / localized kernelised ETKF on the B200 engine.

Reference: pytassim/interface/ketkf.py:32-123, interface/lketkf.py:40-110, core/ketkf.py:28-100.  With
``LinearKernel`` (the reference's default) ``KETKFModule`` reproduces ``ETKFModule`` on centred perturbations (checked with
the reference's own modules: 7e-16), so the device path is the LETKF / ETKF one unchanged.  Every other kernel of
:mod:`pytassim_b200.kernels` (and their ``+``, ``*``, ``**`` compositions) adds one pass between the Gram and the solve
kernels that turns the augmented Gram into the double-centred kernel matrix (csrc/kernelise.cuh).
"""
from .etkf import ETKF
from .letkf import LETKF
from ..kernels import BaseKernel, LinearKernel

__all__ = ['KETKF', 'LKETKF']


def _check_kernel(kernel):
    if not isinstance(kernel, BaseKernel):
        raise NotImplementedError(
            "the B200 engine evaluates kernels on the device from pytassim_b200.kernels descriptors (there is no CPU "
            "fallback for arbitrary torch modules), got {0!r}".format(kernel))
    return kernel


class _KernelMixin(object):
    @property
    def kernel(self):
        return self._kernel

    @kernel.setter
    def kernel(self, new_kernel):                          # interface/ketkf.py:118-123
        self._kernel = _check_kernel(new_kernel).to(dtype=self.dtype, device=self.device)
        self._engines = {}

    def _kernel_key(self):
        kernel = self._kernel
        return () if kernel.is_linear else tuple(kernel.program())

    def _configure_engine(self, engine):
        return engine.set_kernel(self._kernel)

    def _make_core_module(self):
        from ..core import KETKFModule
        return KETKFModule(kernel=self._kernel, inf_factor=self.inf_factor)


class KETKF(_KernelMixin, ETKF):
    def __init__(self, kernel=None, inf_factor=1.0, smoother=False, gpu=False, pre_transform=None, post_transform=None,
                 weight_save_path=None, forward_model=None):
        super().__init__(inf_factor=inf_factor, smoother=smoother, gpu=gpu, pre_transform=pre_transform,
                         post_transform=post_transform, weight_save_path=weight_save_path, forward_model=forward_model)
        self.kernel = LinearKernel() if kernel is None else kernel

    def __str__(self):
        return 'Global KETKF(inf_factor={0}, kernel={1})'.format(str(self.inf_factor.item()), str(self.kernel))

    def __repr__(self):
        return 'KETKF({0},{1})'.format(repr(self.inf_factor.item()), repr(self.kernel))


class LKETKF(_KernelMixin, LETKF):
    def __init__(self, localization=None, kernel=None, inf_factor=1.0, smoother=False, gpu=False, pre_transform=None,
                 post_transform=None, chunksize=10, weight_save_path=None, forward_model=None):
        super().__init__(localization=localization, inf_factor=inf_factor, smoother=smoother, gpu=gpu,
                         pre_transform=pre_transform, post_transform=post_transform, chunksize=chunksize,
                         weight_save_path=weight_save_path, forward_model=forward_model)
        self.kernel = LinearKernel() if kernel is None else kernel

    def __str__(self):
        return 'Localized KETKF(inf_factor={0}, loc={1}, kernel={2})'.format(
            str(self.inf_factor.item()), str(self.localization), str(self.kernel))

    def __repr__(self):
        return 'LKETKF({0},{1},{2})'.format(repr(self.inf_factor.item()), repr(self.localization), repr(self.kernel))
