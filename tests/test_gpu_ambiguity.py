"""The ambiguity protocol of the mask ``use_obs = weights > epsilon`` (localization/gaspari_cohn.py:135; SURVEY.md hard
part 1): pairs whose taper value lies within 1e-13 of epsilon are recorded by the Gram kernel, decided on the host with
the reference's own numpy expression, handed back as overrides, and the blocks whose mask changed are analysed again.
Also: subset neighbour lists and the candidate-run overflow flag."""
import numpy as np
import pytest
import torch

import letkf_oracle as orc
from pytassim_b200.testing import synthetic as syn

pytestmark = pytest.mark.gpu

EPS = 1e-5


def _cutoff_radius():
    """r* with f2(r*) as close to epsilon as bisection in FP64 gets (numpy expression of the reference)."""
    lo, hi = 1.9, 2.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if orc.gc_f2(np.array([mid]))[0] > EPS:
            lo = mid
        else:
            hi = mid
    return lo, hi


def _line_problem(k=12, n_grid=40, seed=3):
    """1-D line, |x_g - x_o| metric, GC radius 10; three observations sit on the cutoff of grid point 5 (just inside, just
    outside, and a few ulps away), the rest are ordinary."""
    rnd = np.random.RandomState(seed)
    c = 10.0
    lo, hi = _cutoff_radius()
    xg = np.arange(n_grid, dtype=np.float64)
    special = np.array([5.0 + lo * c, 5.0 + hi * c, 5.0 - np.nextafter(hi, 3.0) * c, 5.0 + np.nextafter(lo, 0.0) * c])
    xo = np.concatenate([rnd.uniform(-5.0, n_grid + 5.0, size=60), special])
    state = rnd.normal(size=(1, 1, k, n_grid))
    hx = rnd.normal(size=(k, xo.size))
    y = rnd.normal(size=xo.size)
    perts, innov = syn.obs_space_from_hx(hx, y)
    grid_rows = np.stack([np.zeros(n_grid), xg], axis=1)
    obs_rows = np.stack([np.zeros(xo.size), xo], axis=1)
    return dict(state=state, normed_perts=np.ascontiguousarray(perts), normed_obs=innov, grid_rows=grid_rows,
                obs_rows=obs_rows), c


def _engine(data, c, k):
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200.localization.metrics import AbsDistance1D
    eng = LETKFEngine(k, 1, AbsDistance1D(), c, inf_factor=1.05)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
    return eng


def test_pairs_on_the_cutoff_follow_the_reference_rule():
    data, c = _line_problem()
    k = data["state"].shape[2]
    eng = _engine(data, c, k)
    x = torch.as_tensor(data["state"].reshape(1, k, -1)).cuda()
    xa = eng.analyse(x).cpu().numpy().reshape(data["state"].shape)
    assert eng.last_ambiguous["n"] >= 2, eng.last_ambiguous         # the constructed pairs were seen and decided
    ref, _, lists = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                       data["obs_rows"], orc.dist_abs1d, c, inf_factor=1.05, return_lists=True)
    np.testing.assert_allclose(xa, ref, rtol=1e-10, atol=1e-10)
    off, idx, w, amb, namb = eng.neighbour_lists()
    off, idx = off.cpu().numpy(), idx.cpu().numpy()
    assert namb >= 2
    for g in range(len(lists)):
        np.testing.assert_array_equal(idx[off[g]:off[g + 1]], lists[g])
    # a second analysis on the same observations meets the same pairs, already decided: nothing flips
    eng.analyse(x)
    assert eng.last_ambiguous["flipped"] == 0


@pytest.mark.parametrize("host_says_use", [False, True])
def test_a_host_decision_that_differs_is_applied(host_says_use, monkeypatch):
    """Force the host rule to contradict the device for the constructed pairs (as a numpy build with different last bits
    would): the affected block must be analysed again with the host's mask, every other grid point stays bit-identical."""
    data, c = _line_problem(seed=4)
    k = data["state"].shape[2]
    eng = _engine(data, c, k)
    x = torch.as_tensor(data["state"].reshape(1, k, -1)).cuda()
    base = eng.analyse(x, resolve_ambiguous=False).cpu().numpy().reshape(data["state"].shape)
    n, gi, oi, wdev = eng.pending_status()
    assert n >= 2
    from pytassim_b200.localization import gaspari_cohn as gcmod
    real = gcmod.BaseLocalization.host_decision

    def contrary(self, grid_row, obs_rows):
        use, w = real(self, grid_row, obs_rows)
        return np.full_like(use, host_says_use), np.where(host_says_use, np.maximum(w, EPS), w)
    monkeypatch.setattr(gcmod.BaseLocalization, "host_decision", contrary)
    xa = eng.analyse(x).cpu().numpy().reshape(data["state"].shape)
    assert eng.last_ambiguous["n"] == n
    dev_use = wdev > EPS
    assert eng.last_ambiguous["flipped"] == int((dev_use != host_says_use).sum())
    if not host_says_use:
        assert eng.last_ambiguous["flipped"] > 0          # the pair placed just inside the cutoff is in every device mask
    # oracle with the forced mask for the affected grid points
    dist = orc.dist_abs1d
    for g in np.unique(gi):
        d = dist(data["grid_rows"][g], data["obs_rows"])
        luse, lw = orc.gaspari_cohn_localize(d, c, EPS)
        for j in oi[gi == g]:
            luse[j] = host_says_use
            if host_says_use:
                lw[j] = max(lw[j], EPS)
        W = orc.localized_weights(luse, lw, data["normed_perts"], data["normed_obs"], 1.05)
        ref = orc.apply_weights(data["state"][..., [g]], W[None])
        np.testing.assert_allclose(xa[..., [g]], ref, rtol=1e-10, atol=1e-10)
    others = np.setdiff1d(np.arange(data["state"].shape[-1]), np.unique(gi))
    np.testing.assert_array_equal(xa[..., others], base[..., others])
    # the lists reflect the decision as well
    off, idx, w, _, _ = eng.neighbour_lists()
    off, idx = off.cpu().numpy(), idx.cpu().numpy()
    for g, j in zip(gi, oi):
        assert (j in idx[off[g]:off[g + 1]]) == host_says_use


def test_overrides_are_dropped_with_new_observations():
    data, c = _line_problem(seed=5)
    k = data["state"].shape[2]
    eng = _engine(data, c, k)
    x = torch.as_tensor(data["state"].reshape(1, k, -1)).cuda()
    eng.set_overrides({(3, 7): 0.0})
    a = eng.analyse(x, resolve_ambiguous=False).cpu().numpy()
    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])     # clears the list
    b = eng.analyse(x, resolve_ambiguous=False).cpu().numpy()
    ref, _ = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                data["obs_rows"], orc.dist_abs1d, c, inf_factor=1.05)
    np.testing.assert_allclose(b.reshape(ref.shape), ref, rtol=1e-10, atol=1e-10)
    d = dist = orc.dist_abs1d(data["grid_rows"][3], data["obs_rows"])
    if orc.gaspari_cohn_localize(d, c, EPS)[0][7]:
        assert np.abs(a[..., 3] - b[..., 3]).max() > 0          # the override had removed a real local observation


def test_host_path_runs_the_protocol():
    data, c = _line_problem(seed=6)
    k = data["state"].shape[2]
    eng = _engine(data, c, k)
    out = eng.analyse_host(data["state"].reshape(1, k, -1), data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
    assert eng.last_ambiguous["n"] >= 2
    ref, _ = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                data["obs_rows"], orc.dist_abs1d, c, inf_factor=1.05)
    np.testing.assert_allclose(out.reshape(ref.shape), ref, rtol=1e-10, atol=1e-10)


def test_subset_neighbour_lists_match_the_full_lists():
    from pytassim_b200.engine import LETKFEngine
    from pytassim_b200.localization.metrics import HaversineDistance
    data = syn.sphere_latlon(24, 48, 8, 3000, seed=44)
    eng = LETKFEngine(8, 1, HaversineDistance(6371.0), 1000.0)
    eng.set_grid(data["grid_rows"][:, 1:])
    eng.bin_obs(data["obs_rows"][:, 1:], data["normed_perts"], data["normed_obs"])
    off, idx, w, _, _ = eng.neighbour_lists()
    sel = np.array([0, 17, 500, 1151])
    off2, idx2, w2, _, _ = eng.neighbour_lists(subset=sel)
    off, idx, w, off2, idx2, w2 = [t.cpu().numpy() for t in (off, idx, w, off2, idx2, w2)]
    assert off2[-1] == sum(off[g + 1] - off[g] for g in sel)
    for g in sel:
        np.testing.assert_array_equal(idx2[off2[g]:off2[g + 1]], idx[off[g]:off[g + 1]])
        np.testing.assert_array_equal(w2[off2[g]:off2[g + 1]], w[off[g]:off[g + 1]])
