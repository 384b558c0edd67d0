"""The Gram kernels alone (b200da_letkf_gram) against the oracle's localized Gram matrices:
G_g = sum_j w_gj [y_j; d_j][y_j; d_j]^T with w from ``GaspariCohn.localize_obs`` (localization/gaspari_cohn.py:97-136) and
the sqrt(w) gather of interface/wrapper.py:91-97, i.e. C = Y~ Y~^T (core/etkf.py:68) and b = Y~ d~^T (core/etkf.py:72).

FP64 plans run the DMMA kernel (tolerance 1e-12 of max |G|); FP32 plans with k >= 8 run the tcgen05 kernel
(bf16 hi/lo split operands, FP32 accumulation in tensor memory: tolerance 3e-5 of max |G|), below k = 8 the DMMA kernel on
FP32 inputs."""
import numpy as np
import pytest
import torch

import letkf_oracle as orc
from pytassim_b200.testing import synthetic as syn

pytestmark = pytest.mark.gpu


def _metrics():
    from pytassim_b200.localization import metrics
    return metrics


def _oracle_gram(data, dist_func, radius, sel, dtype, taper="gc"):
    yn = data["normed_perts"].astype(dtype).astype(np.float64)
    d = data["normed_obs"].astype(dtype).astype(np.float64)
    aug = np.concatenate([yn, d[None]], axis=0)                        # (k+1, M)
    out = []
    for gi in sel:
        use, w = orc._localize(taper, dist_func(data["grid_rows"][gi], data["obs_rows"]), radius, 1e-5)
        a = aug[:, use] * np.sqrt(w[use])                              # interface/wrapper.py:91-97
        out.append(a @ a.T)                                            # core/utils.py:172
    return np.stack(out)


def _device_gram(data, metric, radius, dtype, taper="gc"):
    from pytassim_b200.engine import LETKFEngine
    k = data["state"].shape[2]
    eng = LETKFEngine(k, 1, metric, radius, inf_factor=1.0, taper=taper, dtype=dtype)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
    g = eng.local_gram().cpu().numpy()
    return eng, g


def _check(got, want, tol):
    k1 = want.shape[-1]
    mask = np.tril(np.ones((k1, k1), dtype=bool))
    mask[k1 - 1, k1 - 1] = False
    scale = np.abs(want).max()
    err = np.abs(got - want)[:, mask].max()
    assert err <= tol * scale, (err, scale)


CASES = [
    # k, generator, metric, radius, oracle dist
    (8, lambda: syn.lorenz96_1d(300, 8, 2, seed=3), lambda m: m.PeriodicDistance1D(300.0), 12.0, lambda: orc.make_dist_periodic1d(300.0)),
    (32, lambda: syn.lorenz96_1d(400, 32, 1, seed=4), lambda m: m.PeriodicDistance1D(400.0), 25.0, lambda: orc.make_dist_periodic1d(400.0)),
    (40, lambda: syn.lorenz96_1d(700, 40, 2, seed=5), lambda m: m.AbsDistance1D(), 40.0, lambda: orc.dist_abs1d),
    (50, lambda: syn.sphere_latlon(30, 60, 50, 6000, seed=6), lambda m: m.HaversineDistance(6371.0), 1000.0, lambda: orc.make_dist_haversine(6371.0)),
    (64, lambda: syn.sphere_latlon(16, 32, 64, 2500, seed=7), lambda m: m.HaversineDistance(6371.0), 1500.0, lambda: orc.make_dist_haversine(6371.0)),
    (100, lambda: syn.lorenz96_1d(150, 100, 1, seed=8), lambda m: m.PeriodicDistance1D(150.0), 9.0, lambda: orc.make_dist_periodic1d(150.0)),
    (128, lambda: syn.lorenz96_1d(140, 128, 1, seed=9), lambda m: m.PeriodicDistance1D(140.0), 8.0, lambda: orc.make_dist_periodic1d(140.0)),
    (5, lambda: syn.lorenz96_1d(90, 5, 1, seed=10), lambda m: m.AbsDistance1D(), 6.0, lambda: orc.dist_abs1d),
    (16, lambda: syn.sphere_latlon(20, 40, 16, 3000, seed=11), lambda m: m.HaversineDistance(6371.0), 1800.0, lambda: orc.make_dist_haversine(6371.0)),
    # the FP64 Gram keeps (k + 1) mod 8 = 1, 2 or 3 trailing rows of [Yn; d] on the DFMA pipe (letkf_kernel.cuh, ER): one case
    # per remainder and warp layout (kt <= 5: one warp per grid point, <= 7: two, <= 10: four, beyond: eight)
    (9, lambda: syn.lorenz96_1d(200, 9, 1, seed=15), lambda m: m.PeriodicDistance1D(200.0), 7.0, lambda: orc.make_dist_periodic1d(200.0)),
    (26, lambda: syn.lorenz96_1d(300, 26, 1, seed=16), lambda m: m.AbsDistance1D(), 11.0, lambda: orc.dist_abs1d),
    (49, lambda: syn.sphere_latlon(20, 40, 49, 3000, seed=17), lambda m: m.HaversineDistance(6371.0), 1500.0, lambda: orc.make_dist_haversine(6371.0)),
    (58, lambda: syn.lorenz96_1d(260, 58, 1, seed=18), lambda m: m.PeriodicDistance1D(260.0), 14.0, lambda: orc.make_dist_periodic1d(260.0)),
    (73, lambda: syn.lorenz96_1d(180, 73, 1, seed=19), lambda m: m.PeriodicDistance1D(180.0), 9.0, lambda: orc.make_dist_periodic1d(180.0)),
    (122, lambda: syn.lorenz96_1d(130, 122, 1, seed=20), lambda m: m.PeriodicDistance1D(130.0), 8.0, lambda: orc.make_dist_periodic1d(130.0)),
]


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 3e-5)])
@pytest.mark.parametrize("case", range(len(CASES)))
def test_local_gram_against_oracle(case, dtype, tol):
    k, gen, metric, radius, dist = CASES[case]
    data = gen()
    n_grid = data["state"].shape[-1]
    eng, got = _device_gram(data, metric(_metrics()), radius, dtype)
    if dtype == torch.float32:
        assert ("tcgen05" in eng.kernel_name) == (k >= 8)
    sel = np.unique(np.linspace(0, n_grid - 1, 40, dtype=np.int64))
    want = _oracle_gram(data, dist(), radius, sel, np.float64 if dtype == torch.float64 else np.float32)
    _check(got[sel], want, tol)


def test_local_gram_gcinf_fp32():
    m = _metrics()
    data = syn.sphere_latlon(20, 40, 36, 4000, seed=12)
    eng, got = _device_gram(data, m.HaversineDistance(6371.0), 1200.0, torch.float32, taper="gcinf")
    sel = np.arange(0, 800, 37)
    want = _oracle_gram(data, orc.make_dist_haversine(6371.0), 1200.0, sel, np.float32, taper="gcinf")
    _check(got[sel], want, 3e-5)


def test_local_gram_empty_and_partial_blocks_fp32():
    """Grid points without any observation in reach get a zero Gram matrix; blocks of fewer than 128 grid points and a
    number of candidates that is not a multiple of the tile size are handled."""
    m = _metrics()
    data = syn.lorenz96_1d(333, 33, 1, seed=14)
    keep = data["obs_rows"][:, 1] < 77
    data["obs_rows"] = data["obs_rows"][keep]
    data["normed_perts"] = np.ascontiguousarray(data["normed_perts"][:, keep]); data["normed_obs"] = data["normed_obs"][keep]
    eng, got = _device_gram(data, m.AbsDistance1D(), 4.0, torch.float32)
    assert np.abs(got[200:]).max() == 0.0
    sel = np.arange(0, 333, 7)
    want = _oracle_gram(data, orc.dist_abs1d, 4.0, sel, np.float32)
    _check(got[sel], want, 3e-5)
