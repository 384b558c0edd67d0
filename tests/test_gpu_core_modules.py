"""The reference's plug-in points (3) and (4) on the device: ``core_module(normed_perts, normed_obs)``,
``assimilation.module`` (numpy bridge, the reference's interface/wrapper.py:29-62) and ``assimilation.localized_module`` (one grid point per
call, interface/wrapper.py:64-98), mirroring tests/unit_tests/core/test_etkf.py, test_ketkf.py, test_ienks.py and
interface/test_letkf.py of the reference.  Expected values: golden vectors from the reference's own modules."""
import numpy as np
import pytest
import torch

import letkf_oracle as orc
from pytassim_b200 import kernels as K
from pytassim_b200.core import ETKFModule, KETKFModule, IEnKSTransformModule, IEnKSBundleModule
from pytassim_b200.interface import ETKF, LETKF, LKETKF, LocalizedIEnKSTransform
from pytassim_b200.localization import GaspariCohn, AbsDistance1D

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-10, atol=1e-10)


def test_etkf_core_module_known_answers(golden):
    """core/test_etkf.py:142-210: P = [[.75,.25],[.25,.75]], w_mean = [0.1,-0.1]; :91-103 no observations -> sqrt(rho) I;
    :47-60 size mismatch -> ValueError."""
    g = golden("core_kat.npz")
    module = ETKFModule(inf_factor=torch.tensor(1.0))
    w = module(torch.as_tensor(g["normed_perts"]), torch.as_tensor(g["normed_obs"]))
    assert isinstance(w, torch.Tensor) and w.is_cuda and w.shape == (2, 2)
    np.testing.assert_allclose(w.cpu().numpy(), g["W"], atol=1e-12)
    np.testing.assert_allclose(w.cpu().numpy(), np.array([[0.1], [-0.1]]) + g["w_perts"], atol=1e-12)
    w0 = ETKFModule(inf_factor=1.21)(torch.zeros((5, 0), dtype=torch.float64), torch.zeros((1, 0), dtype=torch.float64))
    np.testing.assert_allclose(w0.cpu().numpy(), 1.1 * np.eye(5), atol=1e-14)
    with pytest.raises(ValueError):
        module(torch.zeros((3, 4)), torch.zeros((1, 5)))
    gr = golden("core_random.npz")
    for n in range(8):
        w = ETKFModule(float(gr["rho%d" % n]))(torch.as_tensor(gr["Y%d" % n]), torch.as_tensor(gr["d%d" % n]))
        np.testing.assert_allclose(w.cpu().numpy(), gr["W%d" % n], **TOL)


def test_ketkf_and_ienks_core_modules(golden):
    gk = golden("ketkf_kernels.npz")
    module = KETKFModule(kernel=K.RBFKernel(gamma=0.5 / 40), inf_factor=1.1)
    w = module(torch.as_tensor(gk["c0_perts"]), torch.as_tensor(gk["c0_obs"]))
    np.testing.assert_allclose(w.cpu().numpy(), gk["c0_w_rbf"], **TOL)
    assert str(module).startswith("KETKFModule(RBFKernel")
    gi = golden("ienks.npz")
    perts, obs, tau = gi["c1_perts"], gi["c1_obs"], float(gi["c1_tau"])
    w = IEnKSTransformModule(tau=torch.tensor(tau))(torch.eye(10, dtype=torch.float64), torch.as_tensor(perts), torch.as_tensor(obs))
    np.testing.assert_allclose(w.cpu().numpy(), gi["c1_transform_w0"], **TOL)
    eps = float(gi["c1_eps"])
    w = IEnKSBundleModule(epsilon=eps, tau=tau)(torch.eye(10, dtype=torch.float64), torch.as_tensor(perts * eps), torch.as_tensor(obs))
    np.testing.assert_allclose(w.cpu().numpy(), gi["c1_bundle_w0"], **TOL)


def test_module_and_localized_module_per_grid_point(golden):
    """interface/test_letkf.py:106-157 per grid point: ``localized_module(grid_row, perts, innov, obs_info=...)`` of the
    reference returns the (k, k) weights of that grid point as a numpy array."""
    g = golden("fixture_letkf.npz")
    alg = LETKF(localization=GaspariCohn((10.,), AbsDistance1D()))
    w = alg.module(g["a_perts"], g["a_innov"][None])                       # global: all observations, numpy in / out
    assert isinstance(w, np.ndarray) and w.dtype == np.float64
    np.testing.assert_allclose(w, orc.etkf_weights(g["a_perts"], g["a_innov"], 1.0), **TOL)
    np.testing.assert_allclose(ETKF().module(g["a_perts"], g["a_innov"][None]), w, rtol=0, atol=0)
    lm = alg.localized_module
    for j in (0, 7, 19, 39):
        wj = lm(g["a_grid_rows"][j], g["a_perts"], g["a_innov"][None], obs_info=g["a_obs_rows"])
        np.testing.assert_allclose(wj, g["a_weights"][j], **TOL)
    gk = golden("ketkf_kernels.npz")
    lk = LKETKF(localization=GaspariCohn((10.,), AbsDistance1D()), kernel=K.GaussKernel(lengthscale=np.sqrt(20.)), inf_factor=1.1)
    wj = lk.localized_module(g["a_grid_rows"][5], gk["lketkf_perts"], gk["lketkf_innov"][None], obs_info=g["a_obs_rows"])
    np.testing.assert_allclose(wj, gk["lketkf_weights_gauss"][5], **TOL)
    gi = golden("ienks.npz")
    li = LocalizedIEnKSTransform(lambda st, it: (st, st), localization=GaspariCohn((10.,), AbsDistance1D()),
                                 tau=float(gi["l_transform_tau"]))
    wj = li.localized_module(g["a_grid_rows"][11], gi["l_transform_w0"][11], gi["l_perts"], gi["l_innov"][None],
                             obs_info=g["a_obs_rows"], args_to_skip=(0, ))            # interface/lienks.py:109-112
    np.testing.assert_allclose(wj, gi["l_transform_w1"][11], **TOL)
