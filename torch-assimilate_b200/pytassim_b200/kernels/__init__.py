"""Kernel descriptors for the kernelised ETKF (pytassim/kernels/*.py).

Only the linear kernel ``K(x_i, x_j) = x_i^T x_j`` (pytassim/kernels/linear.py:41-63) runs on the device: with it the
kernelised ensemble-space problem of ``KETKFModule`` (pytassim/core/ketkf.py:69-100) is the ETKF's own Gram matrix — the
centring terms vanish because the observation-space perturbations handed over by the interface are centred
(interface/base.py:367-372) — so ``KETKF`` / ``LKETKF`` reuse the Gram and solve kernels unchanged.
Any other kernel raises ``NotImplementedError``: there is no CPU fallback.
"""

__all__ = ["LinearKernel"]


class LinearKernel(object):
    """pytassim/kernels/linear.py:41-63."""

    def __str__(self):
        return 'LinearKernel'

    def __repr__(self):
        return 'Linear'

    def to(self, *args, **kwargs):            # torch.nn.Module protocol used by the kernel setter (interface/ketkf.py:118-123)
        return self
