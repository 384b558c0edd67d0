"""Grid-point sharding of one LETKF analysis over the GPUs of one node (one process per GPU, torch.distributed).

The reference's only parallelism is data-parallel over grid points with the observation arrays handed to every task
(pytassim/interface/letkf.py:121-143: ``state_info.chunk({'grid': chunksize})``, obs arrays one chunk).  The B200
equivalent: rank 0 broadcasts the observation-space arrays once, every rank bins them and analyses a contiguous range
of grid-point blocks, and the analysed columns are all-gathered (NCCL over NVLink on GPUs; the same code runs on gloo /
CPU tensors for the host-logic tests).  Grid points are independent given the observations, so there is no other
data-path collective.

``engine`` is duck-typed: ``n_blocks``, ``block_offset(b)``, ``analyse(x, out=, blocks=(b0, b1))``,
``ienks_step(x, weights, tau=, epsilon=, out=, blocks=)``, ``pack_columns(xa, b0, b1) -> (rows, ncols)`` and
``unpack_columns(packed, b0, b1, xa)``.  An engine with a kernel program (``set_kernel``) shards the same way: the
kernelise pass is per grid point.
"""
import torch
import torch.distributed as dist

__all__ = ["block_range", "ShardedAnalysis", "ShardedETKF"]


def block_range(n_blocks, world_size, rank):
    """Contiguous, balanced split of the block-sorted grid (blocks are spatially ordered and equally sized, so equal
    block counts are equal work for a uniform observation density)."""
    return (n_blocks * rank) // world_size, (n_blocks * (rank + 1)) // world_size


class ShardedAnalysis(object):
    def __init__(self, engine, group=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        nb = engine.n_blocks
        self.ranges = [block_range(nb, self.world, r) for r in range(self.world)]
        self.ncols = [engine.block_offset(b1) - engine.block_offset(b0) for b0, b1 in self.ranges]
        self._send = None
        self._recv = None

    def broadcast_inputs(self, tensors, src=0):
        """Observation-space arrays (and the state) live on ``src``; one broadcast each per analysis."""
        if self.world > 1:
            for t in tensors:
                dist.broadcast(t, src, group=self.group)

    def run(self, x, out):
        """Analyse this rank's blocks of ``x`` (n_slices, k, N) into ``out`` and fill in every other rank's columns."""
        b0, b1 = self.ranges[self.rank]
        self.engine.analyse(x, out=out, blocks=(b0, b1))
        return self._gather_columns(out)

    def run_ienks(self, x, weights, out, tau=1.0, epsilon=None):
        """One localized IEnKS iteration (interface/lienks.py:68-118) on this rank's blocks: ``weights`` is (k, k) or the
        (N, k, k) array of the previous iteration, of which only this rank's grid points are read.  Returns the updated
        (N, k, k) weights — valid for this rank's grid points, which is all the next iteration needs here, so the weights are
        never communicated — and fills ``out`` with the state updated by every rank's weights (all-gathered columns)."""
        b0, b1 = self.ranges[self.rank]
        _, new_weights = self.engine.ienks_step(x, weights, tau=tau, epsilon=epsilon, out=out, blocks=(b0, b1))
        self._gather_columns(out)
        return new_weights

    def _gather_columns(self, out):
        b0, b1 = self.ranges[self.rank]
        if self.world == 1:
            return out
        rows = out.shape[0] * out.shape[1]
        maxc = max(self.ncols)
        if self._send is None or self._send.shape != (rows, maxc) or self._send.device != out.device:
            self._send = torch.zeros((rows, maxc), dtype=out.dtype, device=out.device)
            self._recv = torch.empty((self.world * rows, maxc), dtype=out.dtype, device=out.device)
        packed = self.engine.pack_columns(out, b0, b1)
        self._send[:, :packed.shape[1]] = packed
        dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        recv = self._recv.view(self.world, rows, maxc)
        for r, (rb0, rb1) in enumerate(self.ranges):
            if r != self.rank and self.ncols[r] > 0:
                self.engine.unpack_columns(recv[r, :, :self.ncols[r]].contiguous(), rb0, rb1, out)
        return out


class ShardedETKF(object):
    """Global ETKF (no localization; pytassim/interface/etkf.py:99-120) over the GPUs of one node.

    The reference makes one torch call on the whole observation vector and one einsum over the whole state.  Both shard:
    the augmented Gram ``[Yn; d][Yn; d]^T`` is a sum over observations, the update is independent per grid column.  Rank r
    computes the Gram of its observation range, ONE all-reduce of (k+1)^2 doubles gives every rank the full Gram, every rank
    solves the same k x k problem redundantly (cheaper than broadcasting W from one rank) and updates its own range of state
    columns; ``gather=True`` all-gathers the analysed columns so every rank ends with the whole analysis.

    ``engine`` is duck-typed: ``etkf_gram(yn, d, obs_range=(j0, j1)) -> (k+1, k+1)``,
    ``etkf_weights_from_gram(gram, n_obs_total) -> (k, k)`` and ``apply_weights_cols(x, w, c0, c1, out)``.
    """

    def __init__(self, engine, group=None, align=16):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.align = int(align)

    def ranges(self, n):
        """Contiguous split of n columns into world ranges whose boundaries are multiples of ``align`` (16-byte aligned rows
        for the vectorised stores of the update kernel and whole DMMA fragments for the Gram kernel)."""
        units = (n + self.align - 1) // self.align
        out = []
        for r in range(self.world):
            u0, u1 = block_range(units, self.world, r)
            out.append((min(u0 * self.align, n), min(u1 * self.align, n)))
        return out

    def weights(self, normed_perts, normed_obs):
        """W (k, k), identical on every rank.  normed_perts (k, M) / normed_obs (M,) are the full arrays (every rank reads
        only its own column range)."""
        m = int(normed_obs.shape[-1])
        j0, j1 = self.ranges(m)[self.rank]
        gram = self.engine.etkf_gram(normed_perts, normed_obs, obs_range=(j0, j1))
        if self.world > 1:
            dist.all_reduce(gram, op=dist.ReduceOp.SUM, group=self.group)
        return self.engine.etkf_weights_from_gram(gram, m)

    def run(self, x, normed_perts, normed_obs, out, gather=True):
        """x, out: (n_slices, k, N).  Updates this rank's column range of ``out`` (all columns when ``gather``)."""
        w = self.weights(normed_perts, normed_obs)
        n = int(x.shape[-1])
        cols = self.ranges(n)
        c0, c1 = cols[self.rank]
        self.engine.apply_weights_cols(x, w, c0, c1, out)
        if gather and self.world > 1:
            rows = out.shape[0] * out.shape[1]
            maxc = max(b - a for a, b in cols)
            send = torch.zeros((rows, maxc), dtype=out.dtype, device=out.device)
            send[:, :c1 - c0] = out.view(rows, n)[:, c0:c1]
            recv = torch.empty((self.world * rows, maxc), dtype=out.dtype, device=out.device)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            recv = recv.view(self.world, rows, maxc)
            for r, (a, b) in enumerate(cols):
                if r != self.rank and b > a:
                    out.view(rows, n)[:, a:b] = recv[r, :, :b - a]
        return out
