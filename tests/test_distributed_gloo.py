"""world_size-2 gloo test of the grid-point sharding (pytassim_b200.parallel) on CPU tensors: a stand-in engine that
'analyses' with the CPU oracle exercises block ranges, packing, the padded all-gather and the unpacking."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import letkf_oracle as orc
from pytassim_b200.parallel import ShardedAnalysis, ShardedETKF, block_range
from pytassim_b200.testing import synthetic as syn


class OracleEngine(object):
    """CPU stand-in with the engine's sharding interface; blocks of 8 consecutive grid points in a shuffled order."""

    def __init__(self, data, radius, rho, period):
        self.data, self.radius, self.rho, self.period = data, radius, rho, period
        n = data["state"].shape[-1]
        self.order = np.random.RandomState(5).permutation(n)
        self.offsets = list(range(0, n, 8)) + [n]
        self.n_blocks = len(self.offsets) - 1

    def block_offset(self, b):
        return self.offsets[b]

    def analyse(self, x, out=None, blocks=None):
        b0, b1 = blocks
        sel = self.order[self.offsets[b0]:self.offsets[b1]]
        d = self.data
        ana, _ = orc.letkf_analysis(x.numpy()[None], d["normed_perts"], d["normed_obs"], d["grid_rows"], d["obs_rows"],
                                    orc.make_dist_periodic1d(self.period), self.radius, inf_factor=self.rho, grid_subset=sel)
        out[:, :, torch.as_tensor(sel)] = torch.as_tensor(ana[0])
        return out

    def ienks_step(self, x, weights, tau=1.0, epsilon=None, out=None, blocks=None):
        b0, b1 = blocks
        sel = self.order[self.offsets[b0]:self.offsets[b1]]
        d = self.data
        k, n = x.shape[1], x.shape[2]
        w_in = weights.numpy() if isinstance(weights, torch.Tensor) else np.asarray(weights)
        w_out = torch.full((n, k, k), float("nan"), dtype=torch.float64)
        dist_func = orc.make_dist_periodic1d(self.period)
        for g in sel:
            w_g = w_in if w_in.ndim == 2 else w_in[g]
            w_new = orc.lienks_weights_point(d["grid_rows"][g], w_g, d["normed_perts"], d["normed_obs"][None], d["obs_rows"],
                                             dist_func, (self.radius,), tau, epsilon)
            w_out[g] = torch.as_tensor(w_new)
            out[:, :, g] = torch.as_tensor(orc.apply_weights(x.numpy()[None][..., g:g + 1], w_new[None])[0, ..., 0])
        return out, w_out

    def pack_columns(self, xa, b0, b1, out=None):
        sel = torch.as_tensor(self.order[self.offsets[b0]:self.offsets[b1]])
        packed = xa.reshape(-1, xa.shape[-1])[:, sel].contiguous()
        if out is None:
            return packed
        out.copy_(packed)                     # a strided slot of the all-gather buffer
        return out

    def block_costs(self, counts):
        c = np.asarray(counts, dtype=np.float64)[self.order]
        return np.array([c[self.offsets[b]:self.offsets[b + 1]].sum() for b in range(self.n_blocks)])

    def unpack_columns(self, packed, b0, b1, xa):
        sel = torch.as_tensor(self.order[self.offsets[b0]:self.offsets[b1]])
        xa.view(-1, xa.shape[-1])[:, sel] = packed
        return xa


def _worker(rank, world, port, n_grid, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = syn.lorenz96_1d(n_grid, 6, 2, seed=21)
    x = torch.as_tensor(data["state"][0]).clone()                      # (1, k, N)
    if rank != 0:                                                      # inputs live on rank 0 only
        x.zero_()
        data["normed_perts"] = np.zeros_like(data["normed_perts"]); data["normed_obs"] = np.zeros_like(data["normed_obs"])
    yp, yo = torch.as_tensor(data["normed_perts"]), torch.as_tensor(data["normed_obs"])
    eng = OracleEngine(data, 4.0, 1.1, float(n_grid))
    sh = ShardedAnalysis(eng)
    sh.broadcast_inputs([x, yp, yo])
    data["normed_perts"], data["normed_obs"] = yp.numpy(), yo.numpy()
    out = torch.full_like(x, float("nan"))
    sh.run(x, out)
    ret[rank] = out.numpy()
    dist.destroy_process_group()


def _worker_ienks(rank, world, port, n_grid, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = syn.lorenz96_1d(n_grid, 6, 2, seed=22)
    x = torch.as_tensor(data["state"][0]).clone()
    sh = ShardedAnalysis(OracleEngine(data, 4.0, 1.0, float(n_grid)))
    out = torch.full_like(x, float("nan"))
    weights = torch.eye(6, dtype=torch.float64)
    for _ in range(2):                                                  # the rank's weights stay on the rank between iterations
        weights = sh.run_ienks(x, weights, out, tau=0.8)
    ret[rank] = (out.numpy(), weights.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_ienks_two_ranks_gloo():
    """Two chained localized IEnKS iterations over two ranks: every rank ends with the whole analysis, each rank holds the
    weights of its own grid points only (the other rank's entries are never needed and stay NaN)."""
    n_grid = 43
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_worker_ienks, args=(2, port, n_grid, ret), nprocs=2, join=True)
    data = syn.lorenz96_1d(n_grid, 6, 2, seed=22)
    dist_func = orc.make_dist_periodic1d(float(n_grid))
    ws = np.stack([np.eye(6)] * n_grid)
    for _ in range(2):
        ws = np.stack([orc.lienks_weights_point(data["grid_rows"][g], ws[g], data["normed_perts"], data["normed_obs"][None],
                                                data["obs_rows"], dist_func, (4.0,), 0.8, None) for g in range(n_grid)])
    ref = orc.apply_weights(data["state"], ws)[0]
    owned = np.zeros(n_grid, dtype=int)
    for rank in (0, 1):
        out, weights = ret[rank]
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
        mine = ~np.isnan(weights[:, 0, 0])
        np.testing.assert_allclose(weights[mine], ws[mine], rtol=1e-12, atol=1e-12)
        owned += mine
    assert np.array_equal(owned, np.ones(n_grid, dtype=int))           # every grid point is owned by exactly one rank


def test_block_range_partitions_everything():
    for nb in (1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            rs = [block_range(nb, world, r) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == nb
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1


@pytest.mark.timeout(300)
def test_sharded_analysis_two_ranks_gloo():
    n_grid = 83                                                        # not a multiple of the block size: ragged tail
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_worker, args=(2, port, n_grid, ret), nprocs=2, join=True)
    data = syn.lorenz96_1d(n_grid, 6, 2, seed=21)
    ref, _ = orc.letkf_analysis(data["state"], data["normed_perts"], data["normed_obs"], data["grid_rows"],
                                data["obs_rows"], orc.make_dist_periodic1d(float(n_grid)), 4.0, inf_factor=1.1)
    for rank in (0, 1):
        np.testing.assert_allclose(ret[rank], ref[0], rtol=1e-12, atol=1e-12)


class OracleEtkfEngine(object):
    """CPU stand-in with the engine's sharded-ETKF interface (numpy restatement of core/etkf.py on the summed Gram)."""

    def __init__(self, k, rho):
        self.k, self.rho = k, rho

    def etkf_gram(self, yn, d, obs_range=None):
        j0, j1 = obs_range
        a = np.concatenate([yn.numpy()[:, j0:j1], d.numpy()[None, j0:j1]], axis=0)
        g = np.tril(a @ a.T)
        g[self.k, self.k] = 0.0
        return torch.as_tensor(g)

    def etkf_weights_from_gram(self, gram, n_obs_total):
        k = self.k
        if n_obs_total == 0:
            return torch.as_tensor(np.sqrt(self.rho) * np.eye(k))
        g = gram.numpy()
        c = np.tril(g[:k, :k]) + np.tril(g[:k, :k], -1).T
        b = g[k, :k]
        evals, evects = np.linalg.eigh(c)                                   # core/utils.py:26-61
        evals = np.clip(evals, 0.0, None) + (k - 1) / self.rho
        cov = (evects / evals) @ evects.T                                   # core/etkf.py:70
        w_mean = cov @ b                                                    # core/etkf.py:72
        w_perts = (evects * np.sqrt((k - 1) / evals)) @ evects.T            # core/etkf.py:74-76
        return torch.as_tensor(w_mean[:, None] + w_perts)                   # core/etkf.py:102

    def apply_weights_cols(self, x, w, c0, c1, out):
        xs = x.numpy()[:, :, c0:c1]
        mean = xs.mean(axis=1, keepdims=True)
        out[:, :, c0:c1] = torch.as_tensor(mean + np.einsum('sig,ij->sjg', xs - mean, w.numpy()))
        return out


def _etkf_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(9)
    k, n, m = 7, 53, 37                                                    # ragged against the 16-column alignment
    x = torch.as_tensor(rng.normal(size=(2, k, n)))
    hx = rng.normal(size=(k, m))
    yn = torch.as_tensor(hx - hx.mean(axis=0, keepdims=True))
    d = torch.as_tensor(rng.normal(size=m))
    sh = ShardedETKF(OracleEtkfEngine(k, 1.1))
    out = torch.full_like(x, float("nan"))
    sh.run(x, yn, d, out, gather=True)
    own = torch.full_like(x, float("nan"))
    sh.run(x, yn, d, own, gather=False)
    ret[rank] = (out.numpy(), own.numpy(), sh.ranges(n))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_etkf_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_etkf_worker, args=(2, port, ret), nprocs=2, join=True)
    rng = np.random.RandomState(9)
    k, n, m = 7, 53, 37
    x = rng.normal(size=(2, k, n)); hx = rng.normal(size=(k, m)); d = rng.normal(size=m)
    yn = hx - hx.mean(axis=0, keepdims=True)
    ref, _ = orc.etkf_analysis(x[:, None], yn, d, inf_factor=1.1)           # (n_var, n_time, k, N)
    ref = ref[:, 0]
    for rank in (0, 1):
        out, own, cols = ret[rank]
        np.testing.assert_allclose(out, ref, rtol=1e-11, atol=1e-11)
        c0, c1 = cols[rank]
        np.testing.assert_allclose(own[:, :, c0:c1], ref[:, :, c0:c1], rtol=1e-11, atol=1e-11)
        mask = np.ones(n, bool); mask[c0:c1] = False
        assert np.isnan(own[:, :, mask]).all()                             # other ranks' columns untouched without gather
    assert ret[0][2] == [(0, 32), (32, 53)]


def test_sharded_etkf_ranges_cover_and_align():
    class _E(object):
        pass
    sh = ShardedETKF(_E())
    for world in (1, 2, 3, 8):
        sh.world = world
        for n in (0, 1, 15, 16, 17, 1000, 10_000_000):
            rs = sh.ranges(n)
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert all(a % 16 == 0 for a, _ in rs if a < n)


def _worker_inputs(rank, world, port, ret):
    """InputBuffer: views of one flat buffer travel with one collective (broadcast from the owner, or all-gather of the
    slices every rank filled); then the work split by counts and its feedback re-cut."""
    from pytassim_b200.parallel import InputBuffer
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    specs = [((7, 2), torch.float64), ((3, 7), torch.float32), ((7,), torch.float32), ((1, 3, 11), torch.float64)]
    ref = [torch.arange(int(np.prod(s)), dtype=torch.float64).reshape(s).to(dt) + 10 * i for i, (s, dt) in enumerate(specs)]
    buf = InputBuffer(specs, "cpu", world, rank=rank)
    if rank == 0:
        for v, r in zip(buf.views, ref):
            v.copy_(r)
    else:
        buf.flat.zero_()
    buf.broadcast(0)
    ok_b = all(torch.equal(v, r) for v, r in zip(buf.views, ref))
    # all-gather transport: every rank owns (here: keeps) only its slice of the bytes
    whole = buf.flat.clone()
    buf.flat.zero_()
    buf.slice_of(rank).copy_(whole[rank * (buf.nbytes // world):(rank + 1) * (buf.nbytes // world)])
    buf.allgather(rank)
    ok_a = torch.equal(buf.flat, whole) and all(v.data_ptr() >= buf.flat.data_ptr() for v in buf.views)
    # work split: counts put all the work into the first third of the grid
    data = syn.lorenz96_1d(96, 4, 2, seed=3)
    eng = OracleEngine(data, 4.0, 1.1, 96.0)
    eng.device = "cpu"
    counts = np.zeros(96); counts[eng.order[:32]] = 100.0; counts[eng.order[32:]] = 1.0
    r0 = list(ShardedAnalysis(eng, weights=counts).ranges)
    # uniform counts: an even split; then rank 1 reports twice the time of rank 0: its range must shrink
    sh = ShardedAnalysis(eng, weights=np.ones(96))
    even = list(sh.ranges)
    r1 = list(sh.rebalance(10.0 if rank == 0 else 20.0))
    ret[rank] = (ok_b, ok_a, r0, r1, even)
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_input_buffer_and_rebalance_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    manager = mp.Manager()
    ret = manager.dict()
    mp.spawn(_worker_inputs, args=(2, port, ret), nprocs=2, join=True)
    for rank in (0, 1):
        ok_b, ok_a, r0, r1, even = ret[rank]
        assert ok_b and ok_a
        assert r0 == ret[0][2] and r1 == ret[0][3]                      # every rank uses the same boundaries
        assert r0[0][0] == 0 and r0[-1][1] == 12 and r0[0][1] == r0[1][0]
        # 12 blocks of 8 points, the first 4 carry 800 each, the others 8 each: the cut sits inside the heavy part
        assert r0[0][1] <= 4
        assert even == [(0, 6), (6, 12)]
        assert r1 == [(0, 8), (8, 12)]                                  # rank 1 was twice as slow: it hands two blocks to rank 0
