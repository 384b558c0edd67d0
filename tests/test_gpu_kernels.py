"""Kernelised ETKF (KETKF / LKETKF with non-linear kernels) on the GPU: Gram -> k_kernelise -> ensemble-space solve, through
the C ABI (b200da_plan_set_kernel), against the reference's ``KETKFModule(kernel)`` (tests/golden/ketkf_kernels.npz, generated
from pytassim/core/ketkf.py + pytassim/kernels/*.py by ``oracle/make_golden.py kernels``) and against the oracle restatement
on seeded inputs.  Tolerance: FP64 weights / analysis rtol = atol = 1e-10 (the north star's FP64 bound)."""
import numpy as np
import pytest
import torch

import letkf_oracle as orc
from pytassim_b200 import kernels as K
from pytassim_b200.engine import LETKFEngine
from pytassim_b200.localization import metrics as m
from pytassim_b200.testing import kernel_cases as kc
from pytassim_b200.testing import synthetic as syn

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-10, atol=1e-10)
NAMES = [name for name, _ in kc.KERNEL_CASES]
BUILD = dict(kc.KERNEL_CASES)


@pytest.mark.parametrize("name", NAMES)
def test_global_ketkf_weights_against_reference(golden, name):
    """interface/ketkf.py -> core/ketkf.py:69-100: one weight matrix from all observations, every problem size of the golden
    file (k = 24 and 40 take the Gram variant with the innovation row inside the tiles)."""
    g = golden("ketkf_kernels.npz")
    for i, (k, p, rho) in enumerate(kc.PROBLEM_SIZES):
        kernel = BUILD[name](K, p)
        eng = LETKFEngine(k, 1, m.AbsDistance1D(), 1.0, inf_factor=rho).set_kernel(kernel)
        w = eng.etkf_weights(g["c%d_perts" % i], g["c%d_obs" % i].reshape(-1)).cpu().numpy()
        np.testing.assert_allclose(w, g["c%d_w_%s" % (i, name)], err_msg="%s k=%d" % (name, k), **TOL)
        # the observation-sharded entry points (Gram with d d^T -> weights) give the same matrix
        gram = eng.etkf_gram(g["c%d_perts" % i], g["c%d_obs" % i].reshape(-1))
        aug = np.concatenate([g["c%d_perts" % i], g["c%d_obs" % i]], axis=0)
        assert abs(float(gram[k, k]) - float(aug[k] @ aug[k])) <= 1e-12 * float(aug[k] @ aug[k])
        w2 = eng.etkf_weights_from_gram(gram, p).cpu().numpy()
        np.testing.assert_allclose(w2, w, rtol=1e-12, atol=1e-12)
        # no observations: inflated prior weights (core/etkf.py:91-95)
        w0 = eng.etkf_weights(np.zeros((k, 0)), np.zeros(0)).cpu().numpy()
        np.testing.assert_allclose(w0, np.eye(k) * np.sqrt(rho), rtol=0, atol=1e-14)


@pytest.mark.parametrize("name", NAMES)
def test_localized_ketkf_fixture_against_reference(golden, name):
    """interface/lketkf.py:84-110 on the reference fixtures (GaspariCohn((10.,), |grid - obs|), first time slice): weights of
    all 40 grid points and the analysis against wrapper_localization(wrapper_bridge(KETKFModule(kernel)))."""
    g = golden("ketkf_kernels.npz")
    state = g["lketkf_state"]
    eng = LETKFEngine(10, 2, m.AbsDistance1D(), 10.0, inf_factor=1.1).set_kernel(BUILD[name](K, 20))
    eng.set_grid(g["lketkf_grid"][:, None])
    eng.bin_obs(g["lketkf_obs_grid"][:, None], g["lketkf_perts"], g["lketkf_innov"])
    xa, w = eng.analyse(torch.as_tensor(state.reshape(2, 10, 40)).cuda(), return_weights=True)
    np.testing.assert_allclose(w.cpu().numpy(), g["lketkf_weights_" + name], **TOL)
    np.testing.assert_allclose(xa.cpu().numpy().reshape(state.shape), g["lketkf_analysis_" + name], **TOL)


@pytest.mark.parametrize("name,k,dtype", [("rbf", 50, torch.float64), ("gauss_scale_diag", 40, torch.float64),
                                         ("tanh", 40, torch.float64), ("poly3", 24, torch.float64),
                                         ("rational", 50, torch.float32), ("rbf", 16, torch.float32)])
def test_localized_ketkf_ring_against_oracle(name, k, dtype):
    """Lorenz-96 ring (cfg2 shape, 2000 grid points, every 2nd observed, periodic distance, c = 20): LKETKF analysis against the
    oracle on a subset of grid points, grid points without observations included (prior weights).  FP32 plans store the
    arrays in FP32 and run the FP64 DMMA Gram (kernelised plans never take the tcgen05 Gram); tolerance 1e-4 there."""
    n = 2000
    data = syn.lorenz96_1d(n, k, 2, seed=7)
    # drop the observations around one stretch of the ring so that some grid points see none
    keep = ~((data["obs_rows"][:, 1] > 900) & (data["obs_rows"][:, 1] < 1100))
    obs_rows, perts, innov = data["obs_rows"][keep], data["normed_perts"][:, keep], data["normed_obs"][keep]
    npd = np.float64 if dtype == torch.float64 else np.float32
    perts, innov, state = perts.astype(npd), innov.astype(npd), data["state"].astype(npd)
    kernel = BUILD[name](K, 38)
    eng = LETKFEngine(k, 1, m.PeriodicDistance1D(float(n)), 20.0, inf_factor=1.1, dtype=dtype).set_kernel(kernel)
    assert "kernelise" in eng.kernel_name and "tcgen05" not in eng.kernel_name
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(obs_rows[:, 1:], perts, innov)
    xa = eng.analyse(torch.as_tensor(state.reshape(1, k, n)).cuda()).cpu().numpy().astype(np.float64)
    sel = np.concatenate([np.arange(0, n, 37), np.arange(980, 1020)])
    okern = BUILD[name](orc, 38)
    p64, i64, s64 = perts.astype(np.float64), innov.astype(np.float64), state.astype(np.float64)
    ws = np.stack([orc.lketkf_weights_point(data["grid_rows"][j], p64, i64[None], obs_rows, orc.make_dist_periodic1d(float(n)),
                                            (20.,), okern, inf_factor=1.1) for j in sel])
    ref = orc.apply_weights(s64[..., sel], ws)
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    np.testing.assert_allclose(xa.reshape(1, 1, k, n)[..., sel], ref, rtol=tol, atol=tol * max(1.0, np.abs(ref).max()))
    prior = orc.apply_weights(s64[..., 1000:1001], np.eye(k)[None] * np.sqrt(1.1))      # grid point 1000 has no local observation
    np.testing.assert_allclose(xa.reshape(1, 1, k, n)[..., 1000:1001], prior, rtol=tol, atol=tol)


def test_kernel_program_validation_and_reset():
    """b200da_plan_set_kernel rejects malformed programs; an empty program returns to the plain ETKF path."""
    import ctypes
    from pytassim_b200 import _cabi
    eng = LETKFEngine(10, 1, m.AbsDistance1D(), 1.0)
    lib = eng.lib

    def set_prog(ops):
        n = len(ops)
        return lib.b200da_plan_set_kernel(eng._plan, n, (ctypes.c_int * max(n, 1))(*ops), (ctypes.c_double * max(n, 1))(*([1.0] * n)),
                                          (ctypes.c_double * max(n, 1))(*([1.0] * n)))
    assert set_prog([_cabi.KOP_ADD]) == _cabi.ERR_INVALID                               # nothing to pop
    assert set_prog([_cabi.KOP_LINEAR, _cabi.KOP_LINEAR]) == _cabi.ERR_INVALID          # two values left
    assert set_prog([99]) == _cabi.ERR_UNSUPPORTED
    assert set_prog([_cabi.KOP_LINEAR] * 9 + [_cabi.KOP_ADD] * 7) == _cabi.ERR_UNSUPPORTED   # stack deeper than 8
    assert set_prog([_cabi.KOP_LINEAR] * 17) == _cabi.ERR_INVALID                        # more than B200DA_MAX_KERNEL_OPS
    rng = np.random.RandomState(5)
    hx = rng.normal(size=(10, 30)); perts = hx - hx.mean(0); obs = rng.normal(size=30)
    w_etkf = eng.etkf_weights(perts, obs).cpu().numpy()
    eng.set_kernel(K.LinearKernel() + K.ScaleKernel(0.))          # a program that is the linear kernel: same weights
    np.testing.assert_allclose(eng.etkf_weights(perts, obs).cpu().numpy(), w_etkf, rtol=1e-11, atol=1e-12)
    assert eng.kernel_name.endswith("+kernelise")
    eng.set_kernel(K.TanhKernel())                                # not positive semi-definite: the plan switches to the Jacobi solver
    eng.set_kernel(None)                                          # ... and back to the default with the plain ETKF
    assert "kernelise" not in eng.kernel_name
    assert np.array_equal(eng.etkf_weights(perts, obs).cpu().numpy(), w_etkf)
    with pytest.raises(NotImplementedError):
        LETKFEngine(128, 1, m.AbsDistance1D(), 1.0).set_kernel(K.RBFKernel())           # multiples of 8 up to k = 120


def test_kernelised_interface_classes(golden):
    """KETKF / LKETKF ``assimilate`` with an RBF kernel on the reference fixtures against the oracle through the same glue
    (observation-space prep, localization, weights, update)."""
    from pytassim_b200.interface import KETKF, LKETKF
    from pytassim_b200.localization import GaspariCohn, AbsDistance1D
    from test_host_logic import _fixture_objects
    g, state, obs = _fixture_objects(golden)
    gk = golden("ketkf_kernels.npz")
    st0, ob0 = state.isel(time=[0]), obs.isel(time=[0])
    for name in ("rbf", "tanh"):
        alg = LKETKF(localization=GaspariCohn((10.,), AbsDistance1D()), kernel=BUILD[name](K, 20), inf_factor=1.1)
        np.testing.assert_allclose(alg.assimilate(st0, ob0).values, gk["lketkf_analysis_" + name], **TOL)
    # global KETKF: the fixture's observation-space variables through the oracle
    alg = KETKF(kernel=BUILD["rational"](K, 20), inf_factor=1.1)
    ana = alg.assimilate(st0, ob0).values
    w = orc.ketkf_weights(gk["lketkf_perts"], gk["lketkf_innov"][None], 1.1, BUILD["rational"](orc, 20))
    np.testing.assert_allclose(ana, orc.apply_weights(gk["lketkf_state"], w), **TOL)


def test_linear_program_is_letkf_at_scale():
    """Size-independent identity at the cfg2 size (N = 100 000, k = 40): the kernelise pass with a program that is the linear
    kernel (centred perturbations: the centring terms vanish, core/ketkf.py:80-92) reproduces the LETKF weights of every grid
    point; k = 40 also switches the Gram to the variant with the innovation row inside the tiles."""
    n, k = 100_000, 40
    data = syn.lorenz96_1d(n, k, 2, seed=42)
    x = torch.as_tensor(data["state"].reshape(1, k, n)).cuda()
    res = []
    for kernel in (None, K.LinearKernel() + K.ScaleKernel(0.)):
        eng = LETKFEngine(k, 1, m.PeriodicDistance1D(float(n)), 20.0, inf_factor=1.1)
        if kernel is not None:
            eng.set_kernel(kernel)
        eng.set_grid(data["grid_rows"][:, 1:])
        eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
        res.append(eng.analyse(x, return_weights=True))
    (xa0, w0), (xa1, w1) = res
    assert float((w1 - w0).abs().max()) <= 1e-10 * float(w0.abs().max())
    assert float((xa1 - xa0).abs().max()) <= 1e-10 * float(xa0.abs().max())
