"""Kernel configurations shared by the golden-vector generator (reference classes), the oracle (numpy restatement) and the
GPU parity tests (device descriptors).  Every builder takes a namespace that offers the reference's class names with the
reference's constructor signatures (pytassim/kernels/*.py) and the length of the perturbation vectors ``p``."""
import math

__all__ = ["KERNEL_CASES", "PROBLEM_SIZES"]

# (ens_size k, observations p, inflation): k = 24 and 40 are multiples of 8 (Gram variant with the innovation row in the tiles)
PROBLEM_SIZES = [(10, 40, 1.1), (24, 60, 1.0), (40, 38, 1.2), (50, 200, 1.05)]


def _gauss(ns, p):
    return ns.GaussKernel(lengthscale=math.sqrt(p))


def _rbf(ns, p):
    return ns.RBFKernel(gamma=0.5 / p)


def _poly2(ns, p):
    return ns.PolyKernel(degree=2., const=1.)


def _poly3(ns, p):
    return ns.PolyKernel(degree=3., const=0.5) * ns.ScaleKernel(scaling=1. / p)


def _tanh(ns, p):
    return ns.TanhKernel(coeff=1. / p, const=0.1)


def _rational(ns, p):
    return ns.RationalKernel(lengthscale=math.sqrt(p), weighting=2.)


def _gauss_scale_diag(ns, p):
    return ns.GaussKernel(lengthscale=math.sqrt(p)) * ns.ScaleKernel(scaling=2.) + ns.DiagKernel(scaling=0.5)


def _linear_plus_poly(ns, p):
    return ns.LinearKernel() + ns.PolyKernel(degree=2., const=1.) * ns.ScaleKernel(scaling=0.01)


def _power(ns, p):
    return (ns.GaussKernel(lengthscale=math.sqrt(p)) + ns.ScaleKernel(scaling=1.)) ** ns.ScaleKernel(scaling=2.)


KERNEL_CASES = [("gauss", _gauss), ("rbf", _rbf), ("poly2", _poly2), ("poly3", _poly3), ("tanh", _tanh),
                ("rational", _rational), ("gauss_scale_diag", _gauss_scale_diag), ("linear_plus_poly", _linear_plus_poly),
                ("power", _power)]
