"""BASELINE cfg3 at FULL size (1000 x 1000 lat-lon grid, k = 50, 2.5 M observations, p ~ 56.6 k local observations per
grid point): the CUDA path through the C ABI against the oracle on a sample of grid points — index lists bit-exact,
analysis <= 1e-10 (FP64 plan) / <= 1e-4 (FP32 plan) — plus the ambiguity count of the whole analysis.

The reference loop this follows: interface/letkf.py:127-143 -> interface/wrapper.py:86-98 -> core/etkf.py:57-103 on
the same inputs (oracle/letkf_oracle.py restates it; pinned by tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

import letkf_oracle as orc
from pytassim_b200.testing import synthetic as syn

pytestmark = pytest.mark.gpu

NLAT = NLON = 1000
K, M = 50, 2_500_000
RADIUS, RHO = 1000.0, 1.1


def sample_points(eng=None):
    """>= 24 grid points: evenly spaced, both poles, the dateline (lon 0 and the last longitude), and the two grid points
    either side of a block boundary of the engine's own decomposition."""
    n = NLAT * NLON
    sel = list(np.linspace(0, n - 1, 20, dtype=np.int64))
    sel += [0, NLON // 2, n - 1, n - NLON // 2]                                   # rows next to the south / north pole
    sel += [500 * NLON, 500 * NLON + NLON - 1, 250 * NLON, 750 * NLON + NLON - 1]    # either side of the dateline
    if eng is not None:
        order = torch.empty(n, dtype=torch.int32, device="cuda")
        import ctypes
        from pytassim_b200 import _cabi
        _cabi.check(eng.lib.b200da_grid_order(eng._plan, ctypes.c_void_p(order.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
        b = eng.n_blocks // 3
        s = eng.block_offset(b)
        sel += [int(order[s - 1]), int(order[s])]                                 # last point of block b-1, first of block b
    return np.unique(np.asarray(sel, dtype=np.int64))


@pytest.fixture(scope="module")
def cfg3():
    return syn.sphere_latlon(NLAT, NLON, K, M, seed=42)


@pytest.fixture(scope="module")
def cfg3_oracle(cfg3):
    """The oracle on the sampled grid points (~0.3 s each)."""
    sel = sample_points()
    ref, w, lists = orc.letkf_analysis(cfg3["state"], cfg3["normed_perts"], cfg3["normed_obs"], cfg3["grid_rows"],
                                       cfg3["obs_rows"], orc.make_dist_haversine(6371.0), RADIUS, inf_factor=RHO,
                                       grid_subset=sel, return_lists=True)
    return sel, ref, lists


def _engine(dtype):
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200.localization.metrics import HaversineDistance
    return LETKFEngine(K, 1, HaversineDistance(6371.0), RADIUS, inf_factor=RHO, dtype=dtype)


def test_cfg3_full_size_fp64(cfg3, cfg3_oracle):
    sel, ref, lists = cfg3_oracle
    eng = _engine(torch.float64)
    eng.set_grid(cfg3["grid_rows"][:, 1:])
    eng.bin_obs(cfg3["obs_rows"][:, 1:], cfg3["normed_perts"], cfg3["normed_obs"])
    x = torch.as_tensor(cfg3["state"].reshape(1, K, -1)).cuda()
    xa = eng.analyse(x)                       # runs the ambiguity protocol (engine.resolve_ambiguous)
    torch.cuda.synchronize()
    print("cfg3 full size: ambiguous pairs {0}".format(eng.last_ambiguous))
    # the block-boundary points of THIS engine's decomposition are checked too (oracle on the two extra points)
    extra = np.setdiff1d(sample_points(eng), sel)
    if extra.size:
        ref2, _, lists2 = orc.letkf_analysis(cfg3["state"], cfg3["normed_perts"], cfg3["normed_obs"], cfg3["grid_rows"],
                                             cfg3["obs_rows"], orc.make_dist_haversine(6371.0), RADIUS, inf_factor=RHO,
                                             grid_subset=extra, return_lists=True)
        sel = np.concatenate([sel, extra]); ref = np.concatenate([ref, ref2], axis=-1); lists = list(lists) + list(lists2)
    assert sel.size >= 24
    got = xa[..., torch.as_tensor(sel, device="cuda")].cpu().numpy().reshape(ref.shape)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-10)
    # index lists of the sampled grid points: bit-exact
    off, idx, w, amb, _ = eng.neighbour_lists(subset=sel)
    off, idx = off.cpu().numpy(), idx.cpu().numpy()
    p = []
    for n, g in enumerate(sel):
        np.testing.assert_array_equal(idx[off[g]:off[g + 1]], lists[n])
        p.append(off[g + 1] - off[g])
    assert 40_000 < np.mean(p) < 70_000        # the operating point of the headline: p ~ 56.6 k
    assert torch.isfinite(xa).all().item()


def test_cfg3_full_size_fp32(cfg3, cfg3_oracle):
    """FP32 plan (tcgen05 Gram): analysis within 1e-4 of max |reference| on the same sample."""
    sel, ref, _ = cfg3_oracle
    eng = _engine(torch.float32)
    assert "tcgen05" in eng.kernel_name
    eng.set_grid(cfg3["grid_rows"][:, 1:])
    eng.bin_obs(cfg3["obs_rows"][:, 1:], cfg3["normed_perts"].astype(np.float32), cfg3["normed_obs"].astype(np.float32))
    x = torch.as_tensor(cfg3["state"].reshape(1, K, -1), dtype=torch.float32).cuda()
    xa = eng.analyse(x)
    torch.cuda.synchronize()
    got = xa[..., torch.as_tensor(sel, device="cuda")].cpu().numpy().astype(np.float64).reshape(ref.shape)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    print("cfg3 full size FP32 plan: max rel err {0:.3e}".format(err))
    assert err <= 1e-4
    assert torch.isfinite(xa).all().item()
