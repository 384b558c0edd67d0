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
from pytassim_b200.parallel import ShardedAnalysis, block_range
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

    def pack_columns(self, xa, b0, b1):
        sel = torch.as_tensor(self.order[self.offsets[b0]:self.offsets[b1]])
        return xa.reshape(-1, xa.shape[-1])[:, sel].contiguous()

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
