"""Array-level front end of the B200 LETKF / ETKF engine.

This is the layer ``LETKF.update_state`` / ``ETKF.update_state`` (pytassim_b200.interface) call once the
xarray glue has produced plain arrays; it owns device buffers (torch tensors) and forwards raw pointers and the
current CUDA stream to the C ABI (include/b200da.h).  torch is plumbing here: memory, streams, distributed.
"""
import ctypes

import numpy as np
import torch

from . import _cabi

_TAPERS = {"gc": _cabi.TAPER_GC, "gcinf": _cabi.TAPER_GCINF}


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _dev(a, dtype=torch.float64, device="cuda"):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device)


class LETKFEngine(object):
    """One plan = one (ensemble size, state slices, metric, taper, inflation) configuration on one GPU.

    Parameters mirror ``LETKF(localization=GaspariCohn(length_scale, dist_func, epsilon), inf_factor)``
    (pytassim/interface/letkf.py:72-92, pytassim/localization/gaspari_cohn.py:60-69); ``metric`` is one of the
    metric objects of :mod:`pytassim_b200.localization.metrics`.
    """

    def __init__(self, ens_size, n_slices, metric, length_scale, epsilon=1e-5, inf_factor=1.0, taper="gc",
                 dtype=torch.float64, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("pytassim_b200 needs a CUDA device (sm_100); there is no CPU fallback")
        if dtype not in (torch.float64, torch.float32):
            raise NotImplementedError("the engine computes in float64 or float32, got {0}".format(dtype))
        self.dtype = dtype
        self._np_dtype = np.float64 if dtype == torch.float64 else np.float32
        self.lib = _cabi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.k, self.n_slices = int(ens_size), int(n_slices)
        self.metric = metric
        radius = np.atleast_1d(np.asarray(length_scale, dtype=np.float64))
        params = np.asarray(list(metric.params) + [0.0], dtype=np.float64)
        handle = ctypes.c_void_p()
        n_extra = int(getattr(metric, "n_extra", 0))           # ProductDistance: extra |x_g - x_o| rows, one radius each
        if radius.size < 1 + n_extra:
            raise IndexError("the metric returns {0} distance rows but only {1} length scale(s) were given".format(
                1 + n_extra, radius.size))                     # gaspari_cohn.py:127 raises IndexError on radius[i]
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_plan_create(
                ctypes.byref(handle), self.k, self.n_slices, int(metric.n_coord) - n_extra, int(metric.metric_id),
                params.ctypes.data_as(_cabi._dp), len(metric.params),
                radius.ctypes.data_as(_cabi._dp), int(radius.size), float(epsilon), float(inf_factor),
                _cabi.F64 if dtype == torch.float64 else _cabi.F32, _TAPERS[taper]))
        self._plan = handle
        if n_extra:
            ext = np.ascontiguousarray(radius[1:1 + n_extra], dtype=np.float64)
            _cabi.check(self.lib.b200da_plan_set_extra(self._plan, n_extra, ext.ctypes.data_as(_cabi._dp)))
        self.n_grid = 0
        self.n_obs = 0
        self._keep = {}
        self.epsilon = float(epsilon)
        self.taper = taper
        self.radius = radius
        self._gc = None                 # (N, n_coord) grid coordinates as given to set_grid (device)
        self._oc = None                 # (M, n_coord) observation coordinates as given to bin_obs (device or host)
        self._overrides = {}            # (grid index, obs index) -> decided weight, since the last bin_obs
        self.last_ambiguous = dict(n=0, flipped=0)

    def __del__(self):
        plan = getattr(self, "_plan", None)
        if plan is not None and plan.value:
            try:
                self.lib.b200da_plan_destroy(plan)
            except Exception:
                pass
            self._plan = None

    # -- set-up ----------------------------------------------------------------------------------------------
    def set_grid(self, grid_coords):
        """grid_coords: (N, n_coord) — the coordinate columns of ``_extract_state_information``
        (interface/mixin_local.py:50-69) without the time column."""
        gc = _dev(grid_coords, device=self.device)
        if gc.dim() == 1:
            gc = gc[:, None]
        if gc.shape[1] != self.metric.n_coord:
            raise ValueError("metric expects {0} coordinate column(s), got {1}".format(self.metric.n_coord, gc.shape[1]))
        soa = gc.t().contiguous()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_set_grid(self._plan, _ptr(soa), gc.shape[0], _stream()))
        self.n_grid = int(gc.shape[0])
        self.n_blocks = int(self.lib.b200da_num_blocks(self._plan))
        self._gc, self._oc, self._overrides = gc, None, {}
        return self

    def bin_obs(self, obs_coords, normed_perts, normed_obs):
        """obs_coords (M, n_coord); normed_perts (k, M); normed_obs (M,) — interface/base.py:359-379 outputs."""
        yn = _dev(normed_perts, dtype=self.dtype, device=self.device)
        d = _dev(normed_obs, dtype=self.dtype, device=self.device).reshape(-1)
        if yn.dim() != 2 or yn.shape[0] != self.k:
            raise ValueError("normed_perts must be (ens_size, n_obs)")
        if yn.shape[-1] != d.shape[-1]:                          # pytassim/core/base.py:28-38
            raise ValueError('Observational size between ensemble ({0:d}) and observations '
                             '({1:d}) do not match!'.format(yn.shape[-1], d.shape[-1]))
        oc = _dev(obs_coords, device=self.device)
        if oc.dim() == 1:
            oc = oc[:, None]
        if oc.shape[0] != d.shape[0]:
            raise ValueError("obs_coords and normed_obs disagree on the number of observations")
        soa = oc.t().contiguous()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_bin_obs(self._plan, _ptr(soa), _ptr(yn), _ptr(d), d.shape[0], _stream()))
        self.n_obs = int(d.shape[0])
        self._oc, self._overrides = oc, {}
        return self

    def obs_prep(self, ens_obs, observations, variance):
        """Observation-space variables on the device for a diagonal R (interface/base.py:359-379, observation.py:241-245):
        ens_obs (k, M), observations (M,), variance (M,) -> (normed_perts (k, M), normed_obs (M,)) device tensors."""
        hx = _dev(ens_obs, dtype=self.dtype, device=self.device)
        y = _dev(observations, dtype=self.dtype, device=self.device).reshape(-1)
        var = _dev(variance, dtype=self.dtype, device=self.device).reshape(-1)
        if hx.dim() != 2 or hx.shape[0] != self.k:
            raise ValueError("ens_obs must be (ens_size, n_obs)")
        if hx.shape[1] != y.shape[0] or var.shape[0] != y.shape[0]:
            raise ValueError('Observational size between ensemble ({0:d}) and observations '
                             '({1:d}) do not match!'.format(hx.shape[1], y.shape[0]))
        yn = torch.empty_like(hx)
        d = torch.empty_like(y)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_obs_prep(self._plan, _ptr(hx), _ptr(y), _ptr(var), y.shape[0], _ptr(yn), _ptr(d), _stream()))
        return yn, d

    def obs_gather_prep(self, pseudo_state, src_offset, member_stride, observations, variance):
        """``obs_prep`` with a column-selecting observation operator fused in front (obs_ops/lorenz_96/identity.py:88-92,
        obs_ops/base_ops.py:63-75): HX[i, j] = pseudo_state.flat[src_offset[j] + i * member_stride]."""
        xp = _dev(pseudo_state, dtype=self.dtype, device=self.device)
        src = _dev(src_offset, dtype=torch.int64, device=self.device).reshape(-1)
        y = _dev(observations, dtype=self.dtype, device=self.device).reshape(-1)
        var = _dev(variance, dtype=self.dtype, device=self.device).reshape(-1)
        m = int(y.shape[0])
        if src.shape[0] != m or var.shape[0] != m:
            raise ValueError('Observational size between ensemble ({0:d}) and observations '
                             '({1:d}) do not match!'.format(int(src.shape[0]), m))
        if m and (int(src.min()) < 0 or int(src.max()) + (self.k - 1) * int(member_stride) >= xp.numel()):
            raise IndexError("src_offset points outside the pseudo state")
        yn = torch.empty((self.k, m), dtype=self.dtype, device=self.device)
        d = torch.empty_like(y)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_obs_gather_prep(self._plan, _ptr(xp), _ptr(src), int(member_stride), _ptr(y), _ptr(var), m,
                                                        _ptr(yn), _ptr(d), _stream()))
        return yn, d

    # -- hot path ----------------------------------------------------------------------------------------------
    def analyse(self, state, out=None, return_weights=False, blocks=None, count_ambiguous=False, resolve_ambiguous=True):
        """state (n_slices, k, N) on the device -> analysis of the same shape (interface/base.py:257-278).

        ``resolve_ambiguous`` (default): after the launch, pairs whose taper value the device found within 1e-13 of epsilon
        are decided on the host with the reference's numpy expression and, where the decision differs from the device's, the
        affected blocks are analysed again with the host's decision (:meth:`resolve_ambiguous`; one stream synchronisation).
        ``last_ambiguous`` holds the counts."""
        x = _dev(state, dtype=self.dtype, device=self.device).reshape(self.n_slices, self.k, self.n_grid)
        xa = torch.empty_like(x) if out is None else out
        if xa.dtype != self.dtype or not xa.is_contiguous():
            raise ValueError("out must be a contiguous {0} tensor".format(self.dtype))
        w = torch.empty((self.n_grid, self.k, self.k), dtype=self.dtype, device=self.device) if return_weights else None
        amb = torch.zeros(1, dtype=torch.int64, device=self.device) if count_ambiguous else None
        b0, b1 = (0, self.n_blocks) if blocks is None else blocks
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_letkf(self._plan, _ptr(x), _ptr(xa), _ptr(w), b0, b1, _ptr(amb), _stream()))
            if resolve_ambiguous:
                self.resolve_ambiguous(lambda c0, c1: _cabi.check(
                    self.lib.b200da_letkf(self._plan, _ptr(x), _ptr(xa), _ptr(w), c0, c1, None, _stream())), (b0, b1))
        res = [xa]
        if return_weights:
            res.append(w)
        if count_ambiguous:
            res.append(int(amb.item()))
        return res[0] if len(res) == 1 else tuple(res)

    def ienks_step(self, state, weights, tau=1.0, epsilon=None, out=None, blocks=None):
        """One localized IEnKS iteration (interface/lienks.py:68-118 -> core/ienks.py:134-151) and the update with the new
        weights (interface/base.py:257-278).  ``weights``: (k, k) (the prior identity of the first iteration) or (N, k, k);
        ``epsilon`` None: transform variant, else the bundle variant.  Returns (analysis (n_slices, k, N), weights (N, k, k))."""
        x = _dev(state, dtype=self.dtype, device=self.device).reshape(self.n_slices, self.k, self.n_grid)
        w_in = _dev(weights, dtype=self.dtype, device=self.device)
        per_grid = 1 if w_in.dim() == 3 else 0
        if tuple(w_in.shape[-2:]) != (self.k, self.k) or (per_grid and w_in.shape[0] != self.n_grid):
            raise ValueError("weights must be (ens_size, ens_size) or (n_grid, ens_size, ens_size)")
        xa = torch.empty_like(x) if out is None else out
        w_out = torch.empty((self.n_grid, self.k, self.k), dtype=self.dtype, device=self.device)
        b0, b1 = (0, self.n_blocks) if blocks is None else blocks
        with torch.cuda.device(self.device):
            def launch(c0, c1):
                _cabi.check(self.lib.b200da_letkf_ienks(self._plan, _ptr(x), _ptr(xa), _ptr(w_in), per_grid, _ptr(w_out), float(tau),
                                                        -1.0 if epsilon is None else float(epsilon), c0, c1, _stream()))
            launch(b0, b1)
            self.resolve_ambiguous(launch, (b0, b1))
        return xa, w_out

    def ienks_weights(self, weights, normed_perts, normed_obs, tau=1.0, epsilon=None):
        """``IEnKSTransformModule.forward`` / ``IEnKSBundleModule.forward`` on the device (core/ienks.py:134-174): one global
        (k, k) weight update from all observations."""
        yn = _dev(normed_perts, dtype=self.dtype, device=self.device)
        d = _dev(normed_obs, dtype=self.dtype, device=self.device).reshape(-1)
        if yn.shape[-1] != d.shape[-1]:
            raise ValueError('Observational size between ensemble ({0:d}) and observations '
                             '({1:d}) do not match!'.format(yn.shape[-1], d.shape[-1]))
        w_in = _dev(weights, dtype=self.dtype, device=self.device)
        if tuple(w_in.shape) != (self.k, self.k):
            raise ValueError("weights must be (ens_size, ens_size)")
        w_out = torch.empty_like(w_in)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_etkf_ienks_weights(self._plan, _ptr(yn), _ptr(d), d.shape[0], _ptr(w_in), float(tau),
                                                           -1.0 if epsilon is None else float(epsilon), _ptr(w_out), _stream()))
        return w_out

    def local_gram(self, blocks=None):
        """(N, k+1, k+1) FP64: the localization-weighted augmented Gram matrix of every grid point (lower triangle),
        i.e. the output of the Gram kernel alone (parity hook; core/etkf.py:68,72 after interface/wrapper.py:91-97)."""
        out = torch.zeros((self.n_grid, self.k + 1, self.k + 1), dtype=torch.float64, device=self.device)
        b0, b1 = (0, self.n_blocks) if blocks is None else blocks
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_letkf_gram(self._plan, _ptr(out), b0, b1, _stream()))
        return out

    def analyse_host(self, state, obs_coords, normed_perts, normed_obs, out=None):
        """End-to-end call with HOST arrays (numpy or pinned CPU tensors): upload, bin, analyse, download."""
        nd = self._np_dtype

        def host(a):
            if isinstance(a, torch.Tensor):
                if a.dtype != self.dtype or not a.is_contiguous() or a.is_cuda:
                    raise ValueError("host tensors must be contiguous CPU tensors of dtype {0}".format(self.dtype))
                return a
            return np.ascontiguousarray(a, dtype=nd)
        x, yn, d = host(state), host(normed_perts), host(normed_obs)
        oc = obs_coords
        if isinstance(oc, torch.Tensor):
            oc_soa = oc if oc.shape[0] == self.metric.n_coord and oc.dim() == 2 and oc.shape[1] != self.metric.n_coord \
                else oc.reshape(-1, self.metric.n_coord).t().contiguous()
        else:
            oc_soa = np.ascontiguousarray(np.asarray(oc, dtype=np.float64).reshape(-1, self.metric.n_coord).T)
        m = int(d.shape[-1])
        if yn.shape[-1] != m:
            raise ValueError('Observational size between ensemble ({0:d}) and observations '
                             '({1:d}) do not match!'.format(yn.shape[-1], m))
        if out is None:
            out = np.empty(x.shape, dtype=nd) if not isinstance(x, torch.Tensor) else torch.empty_like(x)

        def hp(a):
            return ctypes.c_void_p(a.data_ptr()) if isinstance(a, torch.Tensor) else ctypes.c_void_p(a.ctypes.data)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_letkf_host(self._plan, hp(oc_soa), hp(yn), hp(d), m, hp(x), hp(out), _stream()))
            self.n_obs = m
            self._oc, self._overrides = (oc_soa, "soa"), {}

            def redo(c0, c1):                        # rare: a host decision differs from the device's -> those blocks again
                _cabi.check(self.lib.b200da_letkf_host_blocks(self._plan, hp(out), c0, c1, _stream()))
            self.resolve_ambiguous(redo, (0, self.n_blocks))
        return out

    # -- ambiguity protocol (include/b200da.h; SURVEY.md hard part 1) -------------------------------------------------------
    def pending_status(self):
        """(n_found, grid idx, obs idx, device taper value) of the pairs the Gram kernel met inside the ambiguity band since
        the last call; raises if a kernel had to drop candidate cells (B200DA_ERR_OVERFLOW).  Synchronises the stream."""
        cap = 4096
        gi = np.empty(cap, dtype=np.int64); oi = np.empty(cap, dtype=np.int64); w = np.empty(cap, dtype=np.float64)
        n = ctypes.c_int64(0)
        _cabi.check(self.lib.b200da_pending_status(self._plan, cap, gi.ctypes.data_as(_cabi._vp), oi.ctypes.data_as(_cabi._vp),
                                                   w.ctypes.data_as(_cabi._vp), ctypes.byref(n), _stream()))
        nf = min(int(n.value), cap)
        return int(n.value), gi[:nf], oi[:nf], w[:nf]

    def _host_rows(self, gi, oi):
        """[t, coords...] rows of the given grid points / observations for the host decision."""
        idx_g = torch.as_tensor(gi, device=self._gc.device)
        grid = self._gc[idx_g].cpu().numpy()
        if isinstance(self._oc, tuple):                       # (n_coord, M) host array of analyse_host
            soa = self._oc[0]
            soa = soa.numpy() if isinstance(soa, torch.Tensor) else soa
            obs = np.ascontiguousarray(soa[:, oi].T)
        else:
            obs = self._oc[torch.as_tensor(oi, device=self._oc.device)].cpu().numpy()
        zg = np.zeros((grid.shape[0], 1)); zo = np.zeros((obs.shape[0], 1))
        return np.concatenate([zg, grid], axis=1), np.concatenate([zo, obs], axis=1)

    def set_overrides(self, pairs):
        """pairs: dict (grid index, obs index) -> weight (0 = not a local observation); replaces the list on the plan."""
        n = len(pairs)
        gi = np.asarray([p[0] for p in pairs], dtype=np.int64)
        oi = np.asarray([p[1] for p in pairs], dtype=np.int64)
        w = np.asarray([pairs[p] for p in pairs], dtype=np.float64)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_plan_set_overrides(self._plan, n, gi.ctypes.data_as(_cabi._vp), oi.ctypes.data_as(_cabi._vp),
                                                           w.ctypes.data_as(_cabi._vp), _stream()))
        self._overrides = dict(pairs)

    def blocks_of_grid(self, grid_idx):
        gi = np.ascontiguousarray(grid_idx, dtype=np.int64)
        out = np.empty(gi.shape[0], dtype=np.int64)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_blocks_of_grid(self._plan, gi.shape[0], gi.ctypes.data_as(_cabi._vp),
                                                       out.ctypes.data_as(_cabi._vp), _stream()))
        return out

    def resolve_ambiguous(self, relaunch, blocks):
        """Second half of the ambiguity protocol, after a launch over ``blocks``: read the pairs the device could not decide,
        evaluate ``weights > epsilon`` for them with the reference's numpy expression on the host
        (localization/gaspari_cohn.py:120-135 via ``BaseLocalization.host_decision``), hand the decisions to the plan and call
        ``relaunch(b, b + 1)`` for every block in which a decision differs from the device's.  Returns the number of pairs."""
        n, gi, oi, wdev = self.pending_status()
        self.last_ambiguous = dict(n=n, flipped=0)
        if n == 0:
            return 0
        if n > gi.shape[0]:
            raise _cabi.B200DAError("{0} (grid point, observation) pairs lie within 1e-13 of epsilon, more than the {1} the "
                                    "engine records: the localization is degenerate for this geometry".format(n, gi.shape[0]))
        from .localization.gaspari_cohn import GaspariCohn, GaspariCohnInf
        loc = (GaspariCohnInf if self.taper == "gcinf" else GaspariCohn)(self.radius, self.metric, self.epsilon)
        grid_rows, obs_rows = self._host_rows(gi, oi)
        pairs = dict(self._overrides)
        flipped = []
        for g in np.unique(gi):
            sel = np.nonzero(gi == g)[0]
            use, w_ref = loc.host_decision(grid_rows[sel[0]], obs_rows[sel])
            for q, j in enumerate(sel):
                key = (int(g), int(oi[j]))
                dev_use = (pairs[key] > 0.0) if key in pairs else bool(wdev[j] > self.epsilon)
                pairs[key] = float(w_ref[q]) if use[q] else 0.0
                if bool(use[q]) != dev_use:
                    flipped.append(int(g))
        self.set_overrides(pairs)
        self.last_ambiguous = dict(n=n, flipped=len(flipped))
        if flipped:
            b0, b1 = blocks
            for b in np.unique(self.blocks_of_grid(np.asarray(flipped, dtype=np.int64))):
                if b0 <= b < b1:
                    relaunch(int(b), int(b) + 1)
            self.pending_status()                             # the repeated blocks record the same pairs again: drop them
        return n

    # -- neighbour lists -----------------------------------------------------------------------------------------
    def neighbour_lists(self, with_weights=True, subset=None):
        """CSR (offsets int64[N+1], idx int32[nnz], w float64[nnz], ambiguous uint8[nnz], n_ambiguous) of the
        local observations of every grid point, ascending obs index = ``np.nonzero(use_obs)[0]``
        (localization/gaspari_cohn.py:135).  ``subset``: grid indices whose lists are materialised (the others get empty
        segments) — at cfg3 size the full CSR would have 5.7e10 entries."""
        counts = torch.zeros(self.n_grid, dtype=torch.int64, device=self.device)
        namb = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_neighbour_count(self._plan, _ptr(counts), _ptr(namb), _stream()))
            if subset is not None:
                keep = torch.zeros(self.n_grid, dtype=torch.int64, device=self.device)
                keep[torch.as_tensor(np.asarray(subset, dtype=np.int64), device=self.device)] = 1
                counts = counts * keep
            offsets = torch.zeros(self.n_grid + 1, dtype=torch.int64, device=self.device)
            offsets[1:] = torch.cumsum(counts, 0)
            nnz = int(offsets[-1].item())
            idx = torch.empty(max(nnz, 1), dtype=torch.int32, device=self.device)
            w = torch.empty(max(nnz, 1), dtype=torch.float64, device=self.device) if with_weights else None
            amb = torch.zeros(max(nnz, 1), dtype=torch.uint8, device=self.device)
            _cabi.check(self.lib.b200da_neighbour_fill(self._plan, _ptr(offsets), _ptr(idx), _ptr(w), _ptr(amb), _stream()))
        return offsets, idx[:nnz], (w[:nnz] if with_weights else None), amb[:nnz], int(namb.item())

    def ambiguous_pairs(self, capacity=4096):
        """(grid index, obs index, taper value) of every pair whose taper value is within 1e-13 of epsilon."""
        gi = torch.empty(capacity, dtype=torch.int64, device=self.device)
        oi = torch.empty(capacity, dtype=torch.int64, device=self.device)
        w = torch.empty(capacity, dtype=torch.float64, device=self.device)
        n = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_neighbour_ambiguous(self._plan, capacity, _ptr(gi), _ptr(oi), _ptr(w), _ptr(n), _stream()))
        nf = min(int(n.item()), capacity)
        return gi[:nf], oi[:nf], w[:nf], int(n.item())

    # -- global ETKF ------------------------------------------------------------------------------------------
    def etkf_weights(self, normed_perts, normed_obs):
        """``ETKFModule.forward`` on the device (core/etkf.py:79-103) -> W (k, k)."""
        yn = _dev(normed_perts, dtype=self.dtype, device=self.device)
        d = _dev(normed_obs, dtype=self.dtype, device=self.device).reshape(-1)
        if yn.shape[-1] != d.shape[-1]:
            raise ValueError('Observational size between ensemble ({0:d}) and observations '
                             '({1:d}) do not match!'.format(yn.shape[-1], d.shape[-1]))
        w = torch.empty((self.k, self.k), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_etkf_weights(self._plan, _ptr(yn), _ptr(d), d.shape[0], _ptr(w), _stream()))
        return w

    def apply_weights(self, state, weights, out=None):
        """``_apply_weights`` (interface/base.py:257-278): weights (k, k) or (N, k, k)."""
        x = _dev(state, dtype=self.dtype, device=self.device)
        n_grid = x.shape[-1]
        x = x.reshape(self.n_slices, self.k, n_grid)
        w = _dev(weights, dtype=self.dtype, device=self.device)
        per_grid = 1 if w.dim() == 3 else 0
        xa = torch.empty_like(x) if out is None else out
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_apply_weights(self._plan, _ptr(x), _ptr(w), per_grid, n_grid, _ptr(xa), _stream()))
        return xa

    # -- global ETKF, observation- / state-sharded ------------------------------------------------------
    def etkf_gram(self, normed_perts, normed_obs, obs_range=None):
        """Augmented Gram ``[Yn; d][Yn; d]^T`` (core/etkf.py:68,72) of the observation columns ``obs_range = (j0, j1)`` of
        (k, M) / (M,) device arrays (default: all) -> dense (k+1, k+1) FP64, lower triangle filled.  Grams of disjoint
        ranges add up: this is the all-reduce payload of the observation-sharded ETKF."""
        yn = _dev(normed_perts, dtype=self.dtype, device=self.device)
        d = _dev(normed_obs, dtype=self.dtype, device=self.device).reshape(-1)
        if yn.dim() != 2 or yn.shape[0] != self.k:
            raise ValueError("normed_perts must be (ens_size, n_obs)")
        if yn.shape[-1] != d.shape[-1]:
            raise ValueError('Observational size between ensemble ({0:d}) and observations '
                             '({1:d}) do not match!'.format(yn.shape[-1], d.shape[-1]))
        m = int(d.shape[0])
        j0, j1 = (0, m) if obs_range is None else (int(obs_range[0]), int(obs_range[1]))
        if not 0 <= j0 <= j1 <= m:
            raise ValueError("obs_range outside [0, n_obs]")
        gram = torch.empty((self.k + 1, self.k + 1), dtype=torch.float64, device=self.device)
        esz = yn.element_size()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_etkf_gram(self._plan, ctypes.c_void_p(yn.data_ptr() + j0 * esz),
                                                  ctypes.c_void_p(d.data_ptr() + j0 * esz), j1 - j0, m, _ptr(gram), _stream()))
        self._keep["etkf_gram"] = (yn, d)
        return gram

    def etkf_weights_from_gram(self, gram, n_obs_total):
        """W (k, k) from the (summed) augmented Gram: evd -> rev_evd x2 -> w_mean + w_perts (core/etkf.py:57-77,102)."""
        g = _dev(gram, dtype=torch.float64, device=self.device)
        if tuple(g.shape) != (self.k + 1, self.k + 1):
            raise ValueError("gram must be (ens_size + 1, ens_size + 1)")
        w = torch.empty((self.k, self.k), dtype=self.dtype, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_etkf_weights_from_gram(self._plan, _ptr(g), int(n_obs_total), _ptr(w), _stream()))
        return w

    def apply_weights_cols(self, state, weights, col_begin, col_end, out, peers=None):
        """``_apply_weights`` (interface/base.py:257-278) on grid columns [col_begin, col_end) of ``state`` (n_slices, k, N)
        into the same columns of ``out``; the other columns of ``out`` are left untouched.  ``peers``: tensors of ``out``'s shape
        and dtype in the memory of other GPUs (peer-accessible, e.g. symmetric memory) that receive the same columns from the
        kernel itself — the all-gather of the state-sharded global ETKF fused into the update (one (k, k) W only)."""
        x = _dev(state, dtype=self.dtype, device=self.device)
        n_grid = x.shape[-1]
        w = _dev(weights, dtype=self.dtype, device=self.device)
        per_grid = 1 if w.dim() == 3 else 0
        if out.dtype != self.dtype or not out.is_contiguous() or out.numel() != x.numel():
            raise ValueError("out must be a contiguous {0} tensor of the state's shape".format(self.dtype))
        if peers:
            if per_grid:
                raise ValueError("peer destinations need one (ens_size, ens_size) weight matrix")
            for t in peers:
                if t.dtype != self.dtype or not t.is_contiguous() or t.numel() != x.numel():
                    raise ValueError("peer destinations must be contiguous {0} tensors of the state's shape".format(self.dtype))
            arr = (ctypes.c_void_p * len(peers))(*[t.data_ptr() for t in peers])
            with torch.cuda.device(self.device):
                _cabi.check(self.lib.b200da_apply_weights_cols_peers(self._plan, _ptr(x), _ptr(w), int(col_begin), int(col_end), n_grid,
                                                                     _ptr(out), len(peers), arr, _stream()))
            return out
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_apply_weights_cols(self._plan, _ptr(x), _ptr(w), per_grid, int(col_begin), int(col_end),
                                                           n_grid, _ptr(out), _stream()))
        return out

    # -- multi-GPU helpers --------------------------------------------------------------------------------------
    def peer_copy_cols(self, dst, src, col_begin, col_end, stream=None):
        """Columns [col_begin, col_end) of the (rows, N) view of ``src`` into the same place of ``dst`` (a tensor of the same
        shape, possibly in another GPU's memory) with one strided copy-engine transfer on ``stream`` (default: current)."""
        n = src.shape[-1]
        rows = src.numel() // n
        st = ctypes.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
        _cabi.check(self.lib.b200da_peer_copy_cols(_ptr(dst), _ptr(src), rows, n, int(col_begin), int(col_end), src.element_size(), st))

    def block_offset(self, block):
        return int(self.lib.b200da_block_offset(self._plan, int(block)))

    def pack_columns(self, xa, b0, b1, out=None):
        """Columns of blocks [b0, b1) of ``xa`` (n_slices, k, N) -> dense (n_slices * k, ncols) in block-sorted order; ``out``
        may be a view with a row stride (a slot of the all-gather buffer)."""
        ncols = self.block_offset(b1) - self.block_offset(b0)
        rows = self.n_slices * self.k
        packed = torch.empty((rows, ncols), dtype=self.dtype, device=self.device) if out is None else out
        if tuple(packed.shape) != (rows, ncols) or packed.stride(1) != 1 or packed.dtype != self.dtype:
            raise ValueError("pack target must be ({0}, {1}) with unit column stride".format(rows, ncols))
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_pack_columns(self._plan, _ptr(xa), b0, b1, _ptr(packed), int(packed.stride(0)), _stream()))
        return packed

    def unpack_columns(self, packed, b0, b1, xa):
        if packed.stride(1) != 1:
            packed = packed.contiguous()
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_unpack_columns(self._plan, _ptr(packed), b0, b1, _ptr(xa), int(packed.stride(0)), _stream()))
        return xa

    def neighbour_counts(self):
        """(counts int64[N] in original grid order, n_ambiguous): local observations per grid point
        (b200da_neighbour_count; the exact p_g of the FLOP model, and the work measure of the multi-GPU split)."""
        counts = torch.zeros(self.n_grid, dtype=torch.int64, device=self.device)
        namb = torch.zeros(1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_neighbour_count(self._plan, _ptr(counts), _ptr(namb), _stream()))
        return counts, int(namb.item())

    def grid_order(self):
        """int32[N]: block-sorted slot -> original grid index."""
        order = torch.empty(self.n_grid, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self.lib.b200da_grid_order(self._plan, _ptr(order), _stream()))
        return order

    def block_costs(self, counts):
        """Work per block for the multi-GPU split: sum over the block's grid points of (local observations + a constant
        for the k x k solve: 13 k^3 FLOP against 2 k^2 per observation, weighted by the ~4x lower FLOP rate of the solve)."""
        c = torch.as_tensor(counts, device=self.device).double()[self.grid_order().long()] + 26.0 * self.k
        cum = torch.cat([torch.zeros(1, dtype=torch.float64, device=self.device), torch.cumsum(c, 0)])
        host = np.empty(self.n_blocks + 1, dtype=np.int64)
        _cabi.check(self.lib.b200da_block_offsets(self._plan, host.ctypes.data_as(_cabi._vp)))
        offs = torch.as_tensor(host, device=self.device)
        return (cum[offs[1:]] - cum[offs[:-1]]).cpu().numpy()

    # -- introspection ---------------------------------------------------------------------------------------------
    @property
    def kernel_name(self):
        return self.lib.b200da_kernel_name(self._plan).decode()

    @property
    def extra_rows(self):
        """Rows of [Yn; d] the FP64 Gram kernel accumulates with DFMA next to the DMMA tiles (0: all rows in the tiles)."""
        return int(self.lib.b200da_gram_extra_rows(self._plan))

    def enable_timing(self, on=True):
        _cabi.check(self.lib.b200da_enable_timing(self._plan, 1 if on else 0))

    def last_phase_ms(self):
        """(gram_ms, solve_ms) of the last analyse(); call last_kernel_ms() first."""
        return float(self.lib.b200da_last_phase_ms(self._plan, 0)), float(self.lib.b200da_last_phase_ms(self._plan, 1))

    def set_solver(self, name):
        """'newton' (default: tensor-core Newton-Schulz inverse square root) or 'jacobi' (shared-memory Jacobi EVD)."""
        _cabi.check(self.lib.b200da_set_solver(self._plan, {"newton": _cabi.SOLVER_NEWTON_SCHULZ, "jacobi": _cabi.SOLVER_JACOBI}[name]))
        return self

    def set_kernel(self, kernel):
        """Kernelised ETKF (core/ketkf.py:69-100): ``kernel`` is a :mod:`pytassim_b200.kernels` descriptor or None / a
        ``LinearKernel`` for the plain ETKF.  Kernels that are not positive semi-definite switch the plan to the
        eigendecomposition solver, which clamps negative eigenvalues as core/utils.py:58 does.  Call before ``set_grid``."""
        if kernel is None or getattr(kernel, "is_linear", False):
            prog = []
        else:
            if not hasattr(kernel, "program"):
                raise NotImplementedError("the B200 engine needs a pytassim_b200.kernels descriptor, got {0!r}".format(kernel))
            prog = list(kernel.program())
        if len(prog) > _cabi.MAX_KERNEL_OPS:
            raise NotImplementedError("kernel compositions are limited to {0} operations".format(_cabi.MAX_KERNEL_OPS))
        n = len(prog)
        ops = (ctypes.c_int * max(n, 1))(*[int(p[0]) for p in prog])
        p0 = (ctypes.c_double * max(n, 1))(*[float(p[1]) for p in prog])
        p1 = (ctypes.c_double * max(n, 1))(*[float(p[2]) for p in prog])
        _cabi.check(self.lib.b200da_plan_set_kernel(self._plan, n, ops, p0, p1))
        if n and not kernel.positive_semidefinite:
            self.set_solver("jacobi")
            self._kernel_forced_jacobi = True
        elif getattr(self, "_kernel_forced_jacobi", False):       # back to the default solver the previous kernel had replaced
            self.set_solver("newton")
            self._kernel_forced_jacobi = False
        self.kernel = kernel if n else None
        return self

    def collect_stats(self, on=True):
        _cabi.check(self.lib.b200da_collect_stats(self._plan, 1 if on else 0))

    def stats(self):
        out = (ctypes.c_int64 * 16)()
        _cabi.check(self.lib.b200da_get_stats(self._plan, out))
        v = list(out)
        return dict(gram_cycles=v[0], evd_cycles=v[1], sweeps=v[2], evds=v[3], setup_cycles=v[4], tiles=v[5], jacobi_prof=v[8:13])

    def last_kernel_ms(self):
        return float(self.lib.b200da_last_kernel_ms(self._plan))


def launch_count():
    return int(_cabi.load().b200da_launch_count())
