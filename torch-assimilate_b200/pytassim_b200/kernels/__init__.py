"""Kernel descriptors for the kernelised ETKF (pytassim/kernels/*.py).

The reference evaluates ``kernel(x, y)`` with torch on the (localized) observation-space perturbations.  Every kernel below is
an element-wise function of ``x_i . x_j``, ``|x_i|^2`` and ``|x_j|^2``, i.e. of entries of the augmented Gram matrix the Gram
kernels already produce, so ``KETKF`` / ``LKETKF`` keep the ETKF device path: Gram -> ``k_kernelise`` (kernel function, double
centring, centred kernel column of the observations; pytassim/core/ketkf.py:69-100) -> ensemble-space solve -> update
(csrc/kernelise.cuh, ``b200da_plan_set_kernel``).  A kernel object here is a *descriptor*: it holds the reference's constructor
arguments and compiles itself into the postfix program the C ABI takes; it never computes on the host (no CPU fallback).

``OrnsteinUhlenbeckKernel`` and ``PeriodicKernel`` (kernels/orn_uhl.py, kernels/periodic.py) need the L1 distance between
perturbation vectors, which is not a function of the Gram: they raise ``NotImplementedError``.

With ``LinearKernel`` the centring terms vanish (the perturbations handed over by the interface are centred,
interface/base.py:367-372), so the kernelised problem IS the ETKF Gram and no program is set.
"""
from .. import _cabi

__all__ = ["BaseKernel", "CompKernel", "AdditiveKernel", "MultiplicativeKernel", "PowerKernel", "LinearKernel", "GaussKernel",
           "RBFKernel", "PolyKernel", "TanhKernel", "RationalKernel", "ScaleKernel", "DiagKernel", "OrnsteinUhlenbeckKernel",
           "PeriodicKernel"]


def _scalar(value):
    """Parameters may be Python numbers, numpy scalars or 0-dim torch tensors (the reference's defaults are float32 tensors,
    e.g. kernels/rbf.py:57; torch promotes a 0-dim tensor to the dtype of the data, which float() reproduces)."""
    return float(value)


def _scalar32(value):
    """kernels/scale.py:70-72 and kernels/diag.py:66-71 build their matrix from ``torch.ones`` of the default dtype (float32)
    and multiply by the scaling before the result meets the float64 data: the constant the reference really applies is the
    float32-rounded one, and so is ours."""
    import numpy as np
    return float(np.float32(float(value)))


class BaseKernel(object):
    """pytassim/kernels/base_kernels.py:40-60: ``+``, ``*`` and ``**`` compose kernels."""
    #: False for kernels whose (centred) matrix can have negative eigenvalues: those need the eigendecomposition solver,
    #: which clamps them like core/utils.py:58; the Newton-Schulz solver assumes a positive semi-definite matrix.
    positive_semidefinite = True

    def __add__(self, other):
        return AdditiveKernel(self, other)

    def __mul__(self, other):
        return MultiplicativeKernel(self, other)

    def __pow__(self, other):
        return PowerKernel(self, other)

    def to(self, *args, **kwargs):            # torch.nn.Module protocol used by the kernel setter (interface/ketkf.py:118-123)
        return self

    def program(self):
        """Postfix program [(op, p0, p1), ...] for ``b200da_plan_set_kernel``."""
        raise NotImplementedError

    @property
    def is_linear(self):
        return False


class CompKernel(BaseKernel):
    """base_kernels.py:63-83."""
    _op = None
    _sym = '?'

    def __init__(self, kernel_1, kernel_2):
        for kern in (kernel_1, kernel_2):
            if not isinstance(kern, BaseKernel):
                raise NotImplementedError("the B200 engine composes pytassim_b200.kernels objects only, got {0!r}".format(kern))
        self.kernel_1 = kernel_1
        self.kernel_2 = kernel_2

    def __str__(self):
        return "{0:s}{1:s}{2:s}".format(str(self.kernel_1), self._sym, str(self.kernel_2))

    def __repr__(self):
        return "{0:s}{1:s}{2:s}".format(repr(self.kernel_1), self._sym, repr(self.kernel_2))

    @property
    def positive_semidefinite(self):
        return self.kernel_1.positive_semidefinite and self.kernel_2.positive_semidefinite

    def program(self):
        return self.kernel_1.program() + self.kernel_2.program() + [(self._op, 0.0, 0.0)]


class AdditiveKernel(CompKernel):
    """base_kernels.py:70-90: K1 + K2."""
    _op = _cabi.KOP_ADD
    _sym = '+'


class MultiplicativeKernel(CompKernel):
    """base_kernels.py:93-120: K1 * K2 (Schur product: positive semi-definite if both are)."""
    _op = _cabi.KOP_MUL
    _sym = '*'


class PowerKernel(CompKernel):
    """base_kernels.py:123-161: K1 ** K2, element-wise."""
    _op = _cabi.KOP_POW
    _sym = '^'
    positive_semidefinite = False


class LinearKernel(BaseKernel):
    """pytassim/kernels/linear.py:41-63."""

    def __str__(self):
        return 'LinearKernel'

    def __repr__(self):
        return 'Linear'

    @property
    def is_linear(self):
        return True

    def program(self):
        return [(_cabi.KOP_LINEAR, 0.0, 0.0)]


class GaussKernel(BaseKernel):
    """pytassim/kernels/rbf.py GaussKernel: exp(-|x / l - y / l|^2 / 2)."""

    def __init__(self, lengthscale=1.):
        self.lengthscale = lengthscale

    def __str__(self):
        return "GaussKernel(l={0})".format(self.lengthscale)

    def __repr__(self):
        return "GaussKernel"

    def _get_lengthscale(self):
        return self.lengthscale

    def program(self):
        return [(_cabi.KOP_GAUSS, _scalar(self._get_lengthscale()), 0.0)]


class RBFKernel(GaussKernel):
    """pytassim/kernels/rbf.py RBFKernel: exp(-gamma |x - y|^2) through the length scale sqrt(0.5 / gamma), computed in the
    parameter's own type as the reference does (a float32 tensor stays float32 there)."""

    def __init__(self, gamma=0.5):
        self.gamma = gamma

    def __str__(self):
        return "RBFKernel(γ={0})".format(self.gamma)

    def __repr__(self):
        return "RBFKernel"

    def _get_lengthscale(self):
        return (0.5 / self.gamma) ** 0.5


class PolyKernel(BaseKernel):
    """pytassim/kernels/polynomial.py: (x . y + c)^p.  Positive semi-definite for integer p >= 1 and c >= 0."""

    def __init__(self, degree=2., const=1.):
        self.degree = degree
        self.const = const

    def __str__(self):
        return 'PolynomialKernel({0}, {1})'.format(str(self.degree), str(self.const))

    def __repr__(self):
        return 'Polynomial({0}, {1})'.format(repr(self.degree), repr(self.const))

    @property
    def positive_semidefinite(self):
        deg, const = _scalar(self.degree), _scalar(self.const)
        return deg >= 1.0 and deg == int(deg) and const >= 0.0

    def program(self):
        return [(_cabi.KOP_POLY, _scalar(self.degree), _scalar(self.const))]


class TanhKernel(BaseKernel):
    """pytassim/kernels/tanh.py: tanh(alpha x . y + c); not positive semi-definite in general."""
    positive_semidefinite = False

    def __init__(self, coeff=1., const=0.):
        self.coeff = coeff
        self.const = const

    def __str__(self):
        return 'TanhKernel({0}, {1})'.format(str(self.coeff), str(self.const))

    def __repr__(self):
        return 'Tanh({0}, {1})'.format(repr(self.coeff), repr(self.const))

    def program(self):
        return [(_cabi.KOP_TANH, _scalar(self.coeff), _scalar(self.const))]


class RationalKernel(BaseKernel):
    """pytassim/kernels/rational.py: (1 + |x / l - y / l|^2 / (2 a))^(-a)."""

    def __init__(self, lengthscale=1., weighting=1.):
        self.lengthscale = lengthscale
        self.weighting = weighting

    def __str__(self):
        return 'RationalKernel({0}, {1})'.format(str(self.lengthscale), str(self.weighting))

    def __repr__(self):
        return 'Rational({0}, {1})'.format(repr(self.lengthscale), repr(self.weighting))

    def program(self):
        return [(_cabi.KOP_RATIONAL, _scalar(self.lengthscale), _scalar(self.weighting))]


class ScaleKernel(BaseKernel):
    """pytassim/kernels/scale.py: the constant c."""

    def __init__(self, scaling=0.):
        self.scaling = scaling

    def __str__(self):
        return 'ScaleKernel({0})'.format(str(self.scaling))

    def __repr__(self):
        return repr(self.scaling)

    @property
    def positive_semidefinite(self):
        return _scalar(self.scaling) >= 0.0

    def program(self):
        return [(_cabi.KOP_SCALE, _scalar32(self.scaling), 0.0)]


class DiagKernel(BaseKernel):
    """pytassim/kernels/diag.py: c I between the perturbations, zeros against the single observation vector."""

    def __init__(self, scaling=0.):
        self.scaling = scaling

    def __str__(self):
        return 'DiagKernel({0})'.format(str(self.scaling))

    def __repr__(self):
        return 'Diag({0})'.format(repr(self.scaling))

    @property
    def positive_semidefinite(self):
        return _scalar(self.scaling) >= 0.0

    def program(self):
        return [(_cabi.KOP_DIAG, _scalar32(self.scaling), 0.0)]


class _L1Kernel(BaseKernel):
    def __init__(self, *args, **kwargs):
        raise NotImplementedError(
            "{0} needs the L1 distance between perturbation vectors, which is not a function of the Gram matrix; the B200 "
            "engine has no device path (and no CPU fallback) for it".format(type(self).__name__))


class OrnsteinUhlenbeckKernel(_L1Kernel):
    """pytassim/kernels/orn_uhl.py — not available on the device path."""


class PeriodicKernel(_L1Kernel):
    """pytassim/kernels/periodic.py — not available on the device path."""
