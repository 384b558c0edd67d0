"""CPU emulation of the tabulated taper of the FP64 Gram kernels (csrc/common.cuh: TaperTab, taper_tab_eval;
csrc/b200da.cu: build_taper_table): the weight as a function of the bin-space distance u, piecewise degree-5 Chebyshev
interpolants on 128 intervals per unit of r.  The emulation follows the C++ step by step (same nodes, same Chebyshev ->
monomial conversion, same index arithmetic in float64) and is compared with the oracle's Gaspari-Cohn functions
(gaspari_cohn.py:78-95, 172-210): the table must stay below the rounding level of the direct evaluation, far inside the
1e-13 ambiguity band in which the host decides the mask."""
import numpy as np
import pytest

import letkf_oracle as orc

LD = np.longdouble


def _piece(taper, seg, r):
    r2, r3, r4, r5 = r * r, r ** 3, r ** 4, r ** 5
    if taper == "gcinf":
        if seg == 0:
            return -28 * r5 / LD(33) + 8 * r4 / LD(11) + 20 * r3 / LD(11) - 80 * r2 / LD(33) + 1
        if seg == 1:
            return 20 * r5 / LD(33) - 16 * r4 / LD(11) + 100 * r2 / LD(33) - 45 * r / LD(11) + LD(51) / 22 - LD(7) / (44 * r)
        if seg == 2:
            return -4 * r5 / LD(11) + 16 * r4 / LD(11) - 10 * r3 / LD(11) - 100 * r2 / LD(33) + 5 * r - LD(61) / 22 + LD(115) / (132 * r)
        return 4 * r5 / LD(33) - 8 * r4 / LD(11) + 10 * r3 / LD(11) + 80 * r2 / LD(33) - 80 * r / LD(11) + LD(64) / 11 - LD(32) / (33 * r)
    if seg == 0:
        return -LD(0.25) * r5 + LD(0.5) * r4 + LD(0.625) * r3 - LD(5) / 3 * r2 + 1
    return r5 / 12 - LD(0.5) * r4 + LD(0.625) * r3 + LD(5) / 3 * r2 - 5 * r + 4 - LD(2) / 3 / r


def build_table(taper, radius, sphere_r=None):
    """(ub, scale, coef[nseg, nint, 6]) exactly as build_taper_table does."""
    hav = sphere_r is not None
    nseg = 4 if taper == "gcinf" else 2
    nint = 256 // nseg
    dr = LD(2) / nseg
    u_of_r = (lambda r: 2 * np.sin(LD(0.5) * r * LD(radius) / LD(sphere_r))) if hav else (lambda r: r * LD(radius))
    r_of_u = (lambda u: 2 * LD(sphere_r) * np.arcsin(LD(0.5) * u) / LD(radius)) if hav else (lambda u: u / LD(radius))
    ub = np.array([np.float64(u_of_r(dr * s)) for s in range(nseg + 1)])
    scale = np.zeros(nseg)
    coef = np.zeros((nseg, nint, 6))
    j = np.arange(6)
    pi = LD("3.14159265358979323846264338327950288")
    tn = np.cos((2 * j + 1).astype(LD) * pi / 12)
    for s in range(nseg):
        ua, ue = LD(ub[s]), LD(ub[s + 1])
        h = (ue - ua) / nint
        scale[s] = np.float64(1 / h)
        inv = 1 / LD(scale[s])
        for i in range(nint):
            mid, half = ua + (i + LD(0.5)) * inv, LD(0.5) * inv
            r = np.maximum(r_of_u(mid + half * tn), LD(1e-30))
            f = _piece(taper, s, r)
            a = np.array([(f * np.cos(kk * (2 * j + 1).astype(LD) * pi / 12)).sum() * (LD(1) / 6 if kk == 0 else LD(2) / 6)
                          for kk in range(6)], dtype=LD)
            coef[s, i] = [a[0] - a[2] + a[4], a[1] - 3 * a[3] + 5 * a[5], 2 * a[2] - 8 * a[4], 4 * a[3] - 20 * a[5], 8 * a[4],
                          16 * a[5]]
    return ub, scale, coef


def eval_table(tab, u):
    """taper_tab_eval in float64, vectorised."""
    ub, scale, coef = tab
    nseg, nint = coef.shape[0], coef.shape[1]
    u_in = np.asarray(u, dtype=np.float64)
    inside = u_in < ub[nseg]                      # the device returns 0 before it computes an index (also for NaN)
    u = np.where(inside, u_in, 0.0)
    s = np.zeros(u.shape, dtype=np.int64)
    for i in range(1, nseg):
        s[u >= ub[i]] = i
    x = (u - ub[s]) * scale[s]
    i = np.minimum(np.floor(x).astype(np.int64), nint - 1)
    t = 2.0 * (x - i) - 1.0
    c = coef[s, i]
    w = c[:, 5]
    for kk in (4, 3, 2, 1, 0):
        w = t * w + c[:, kk]
    return np.where(inside, w, 0.0)


@pytest.mark.parametrize("taper", ["gc", "gcinf"])
@pytest.mark.parametrize("geom", [("haversine", 1000.0, 6371.0), ("haversine", 3000.0, 6371.0), ("line", 20.0, None),
                                  ("line", 5.0, None), ("haversine", 0.3, 1.0)])
def test_table_reproduces_the_taper(taper, geom):
    kind, radius, sphere_r = geom
    tab = build_table(taper, radius, sphere_r)
    rng = np.random.default_rng(7)
    # distances in the metric's units: dense sample of (0, 2 radius), clustered near 0, the breakpoints and the end of the support
    r = np.concatenate([rng.uniform(0.0, 2.0, 200_000), rng.uniform(0.0, 1e-3, 2000), 1.0 + rng.uniform(-1e-6, 1e-6, 2000),
                        2.0 - rng.uniform(0.0, 1e-3, 2000), np.linspace(0.0, 2.0, 513)[:-1]])
    dist = r * radius
    if sphere_r is not None:
        u = 2.0 * np.sin(0.5 * dist / sphere_r)                # chord of the unit sphere (bin space)
        dist = 2.0 * sphere_r * np.arcsin(0.5 * u)              # the distance the direct evaluation sees for this chord
    else:
        u = dist
    got = eval_table(tab, u)
    # (1) against the taper in extended precision: the table is at least as accurate as the float64 formulas
    dist_l = 2 * LD(sphere_r) * np.arcsin(LD(0.5) * u.astype(LD)) if sphere_r is not None else dist.astype(LD)
    r_l = dist_l / LD(radius)
    nseg = 4 if taper == "gcinf" else 2
    seg = np.minimum((r_l * nseg / 2).astype(np.int64), nseg - 1)
    exact = np.zeros(r_l.shape, dtype=LD)
    for sg in range(nseg):
        m = seg == sg
        exact[m] = _piece(taper, sg, np.maximum(r_l[m], LD(1e-30)))
    assert np.abs(got - exact).max() < (5e-16 if taper == "gc" else 4e-15)
    # (2) against the oracle's float64 evaluation (gaspari_cohn.py:124-134, unmasked): within its own rounding (terms up to 10)
    loc = orc.gaspari_cohn_inf_localize if taper == "gcinf" else orc.gaspari_cohn_localize
    ref = loc(dist, radius, epsilon=0.0)[1]
    assert np.abs(got - ref).max() < 2e-14


def test_table_is_zero_beyond_the_support_and_for_nan():
    tab = build_table("gc", 20.0)
    assert eval_table(tab, [40.0, 41.0, 1e300]).tolist() == [0.0, 0.0, 0.0]
    assert eval_table(tab, [np.nan])[0] == 0.0
