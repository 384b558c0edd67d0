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

__all__ = ["block_range", "balanced_ranges", "InputBuffer", "ShardedAnalysis", "ShardedETKF"]


def block_range(n_blocks, world_size, rank):
    """Contiguous split of the block-sorted grid into equal block counts (equal work for a uniform observation density)."""
    return (n_blocks * rank) // world_size, (n_blocks * (rank + 1)) // world_size


def balanced_ranges(block_cost, world_size):
    """Contiguous block ranges of (nearly) equal summed cost: ``block_cost`` is a 1-D array of per-block work (local
    observation pairs + a constant per grid point for the solve).  Boundaries sit where the cumulative cost crosses
    r / world of the total."""
    import numpy as np
    cost = np.asarray(block_cost, dtype=np.float64)
    nb = cost.shape[0]
    cum = np.concatenate([[0.0], np.cumsum(cost)])
    if not cum[-1] > 0.0:
        return [block_range(nb, world_size, r) for r in range(world_size)]
    bounds = [0]
    for r in range(1, world_size):
        b = int(np.searchsorted(cum, cum[-1] * r / world_size, side="left"))
        bounds.append(min(max(b, bounds[-1]), nb))
    bounds.append(nb)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


class InputBuffer(object):
    """The per-analysis inputs (observation coordinates, Yn, d, state) as views of ONE flat device buffer, so that they move
    between ranks with one collective: ``broadcast()`` when rank 0 owns them, ``allgather()`` when every rank has filled its
    own ``slice_of(rank)`` of the flat bytes (each rank uploads 1 / world of the bytes over its own PCIe link; the pieces
    travel over NVLink).  specs: list of (shape, dtype)."""

    ALIGN = 256

    def __init__(self, specs, device, world=1, group=None, rank=0, mode="broadcast"):
        """``mode``: how ``broadcast`` moves the bytes from the owning rank: "broadcast" (one NCCL broadcast; 1.46 GB to 7 peers
        in 2.3 ms on an NVSwitch box, tools/diag_bcast.py) or "scatter_allgather" (the owner sends every rank its 1 / world
        slice, an in-place all-gather completes the buffer: measured slower there, 3.9 ms)."""
        self.group, self.world, self.rank, self.mode = group, world, rank, mode
        offs, total = [], 0
        for shape, dtype in specs:
            n = 1
            for v in shape:
                n *= int(v)
            nbytes = n * torch.empty((), dtype=dtype).element_size()
            offs.append((total, nbytes))
            total = (total + nbytes + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        unit = self.ALIGN * world
        total = (max(total, 1) + unit - 1) // unit * unit
        self.flat = torch.empty(total, dtype=torch.uint8, device=device)
        self.views = [self.flat[o:o + nb].view(dtype).view(*shape) for (o, nb), (shape, dtype) in zip(offs, specs)]
        self.nbytes = total

    def slice_of(self, rank):
        per = self.nbytes // self.world
        return self.flat[rank * per:(rank + 1) * per]

    def broadcast(self, src=0):
        if self.world <= 1:
            return
        if self.mode == "scatter_allgather" and self.flat.is_cuda:
            mine = self.slice_of(self.rank)
            dist.scatter(mine, [self.slice_of(r) for r in range(self.world)] if self.rank == src else None, src=src, group=self.group)
            dist.all_gather_into_tensor(self.flat, mine, group=self.group)
        else:
            dist.broadcast(self.flat, src, group=self.group)

    def allgather(self, rank):
        if self.world > 1:
            dist.all_gather_into_tensor(self.flat, self.slice_of(rank), group=self.group)      # in place


class ShardedAnalysis(object):
    def __init__(self, engine, group=None, weights=None):
        """``weights``: optional per-grid-point local-observation counts (1-D tensor / array in ORIGINAL grid order, e.g.
        ``engine.neighbour_counts()``): the block ranges are then balanced by work (sum of counts + a per-point constant for
        the solve) instead of by block count, which removes the rank-max of the Gram kernel (1.2 % at cfg3 on 8 GPUs)."""
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        nb = engine.n_blocks
        self._cost = None
        if weights is not None and self.world > 1:
            import numpy as np
            self._cost = np.asarray(engine.block_costs(weights), dtype=np.float64)
            self._set_ranges(balanced_ranges(self._cost, self.world))
        else:
            self.ranges = [block_range(nb, self.world, r) for r in range(self.world)]
            self.ncols = [engine.block_offset(b1) - engine.block_offset(b0) for b0, b1 in self.ranges]
        self._recv = None

    def _set_ranges(self, ranges):
        rt = torch.as_tensor(ranges, dtype=torch.int64, device=getattr(self.engine, "device", "cpu"))
        dist.broadcast(rt, 0, group=self.group)                     # every rank uses rank 0's boundaries
        self.ranges = [(int(a), int(b)) for a, b in rt.cpu().tolist()]
        self.ncols = [self.engine.block_offset(b1) - self.engine.block_offset(b0) for b0, b1 in self.ranges]

    def rebalance(self, my_ms):
        """Feedback step of the work split: ``my_ms`` is the device time this rank needed for its blocks in the last analysis
        (``engine.last_kernel_ms()``).  The cost of every block of a rank is scaled by that rank's time over the mean time (the
        count-based model misses what depends on the geometry: candidates tested per block, refill passes), and the boundaries
        are cut again.  Same observation network and grid in the next analysis (a cycling assimilation): one or two steps
        remove the rank-max of the Gram kernel.  Collective: every rank calls it."""
        if self.world <= 1 or self._cost is None:
            return self.ranges
        import numpy as np
        t = torch.tensor([float(my_ms)], dtype=torch.float64, device=getattr(self.engine, "device", "cpu"))
        allt = [torch.zeros_like(t) for _ in range(self.world)]
        dist.all_gather(allt, t, group=self.group)
        times = np.asarray([float(v) for v in allt])
        if not np.all(times > 0.0):
            return self.ranges
        mean = times.mean()
        for r, (b0, b1) in enumerate(self.ranges):
            self._cost[b0:b1] *= times[r] / mean
        self._set_ranges(balanced_ranges(self._cost, self.world))
        return self.ranges

    def broadcast_inputs(self, tensors, src=0):
        """Observation-space arrays (and the state) live on ``src``: ONE broadcast when ``tensors`` is an
        :class:`InputBuffer`, else one per tensor."""
        if self.world > 1:
            if isinstance(tensors, InputBuffer):
                tensors.broadcast(src)
                return
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
        """Rank r's analysed columns are packed (block-sorted order) straight into slot r of a (world, rows, maxc) buffer,
        all-gathered IN PLACE, and every other rank's slot is scattered into ``out`` by one kernel each, reading the slot
        with its row stride (no staging copies)."""
        b0, b1 = self.ranges[self.rank]
        if self.world == 1:
            return out
        rows = out.shape[0] * out.shape[1]
        maxc = max(self.ncols)
        if self._recv is None or self._recv.shape != (self.world, rows, maxc) or self._recv.device != out.device \
                or self._recv.dtype != out.dtype:
            self._recv = torch.zeros((self.world, rows, maxc), dtype=out.dtype, device=out.device)
        mine = self._recv[self.rank]
        if self.ncols[self.rank] > 0:
            self.engine.pack_columns(out, b0, b1, out=mine[:, :self.ncols[self.rank]])
        dist.all_gather_into_tensor(self._recv.view(self.world * rows, maxc), mine, group=self.group)
        for r, (rb0, rb1) in enumerate(self.ranges):
            if r != self.rank and self.ncols[r] > 0:
                self.engine.unpack_columns(self._recv[r, :, :self.ncols[r]], rb0, rb1, out)
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
        self._recv = None
        self._peers = None
        self._copy_stream = None
        self.pieces = 8                  # column pieces of the overlapped update + peer copies (symmetric outputs)
        self.fused_stores = False        # True: peer stores from inside the update kernel (measured slower: 16-byte NVLink stores)

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

    def alloc_output(self, shape, dtype, device):
        """An analysis array every rank of the group can write into: symmetric memory (CUDA peer mappings over NVLink,
        ``torch.distributed._symmetric_memory``) when the node offers it, else an ordinary tensor.  With a symmetric output
        ``run(..., gather=True)`` needs no collective after the update: the update kernel stores every analysed tile into the
        output of every rank (``b200da_apply_weights_cols_peers``)."""
        self._peers = None
        if self.world > 1 and torch.device(device).type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm
                group = self.group if self.group is not None else dist.group.WORLD
                out = symm.empty(*tuple(int(v) for v in shape), dtype=dtype, device=device)
                hdl = symm.rendezvous(out, group=group)
                peers = [hdl.get_buffer(r, tuple(shape), dtype) for r in range(self.world) if r != self.rank]
                self._peers = (out.data_ptr(), peers, hdl)
                return out
            except Exception as exc:                          # no NVLink peer access / unsupported build: plain gather path
                self._peers = None
                self.symmetric_error = repr(exc)
        return torch.empty(tuple(shape), dtype=dtype, device=device)

    def run(self, x, normed_perts, normed_obs, out, gather=True):
        """x, out: (n_slices, k, N).  Updates this rank's column range of ``out`` (all columns when ``gather``)."""
        w = self.weights(normed_perts, normed_obs)
        n = int(x.shape[-1])
        cols = self.ranges(n)
        c0, c1 = cols[self.rank]
        peers = getattr(self, "_peers", None)
        if gather and self.world > 1 and peers is not None and peers[0] == out.data_ptr() and w.dim() == 2:
            if self.fused_stores:
                # the update kernel itself stores every tile into every rank's array (b200da_apply_weights_cols_peers)
                self.engine.apply_weights_cols(x, w, c0, c1, out, peers=peers[1])
            else:
                # update in column pieces; behind each piece a copy stream pushes it into every peer's array over NVLink
                # (row-contiguous 16-byte stores of whole lines) while the next piece is computed
                rows = out.shape[0] * out.shape[1]
                flat = out.view(rows, n)
                pflat = [p.view(rows, n) for p in peers[1]]
                main = torch.cuda.current_stream()
                if self._copy_stream is None:                 # one copy stream per peer: the transfers to different GPUs overlap
                    self._copy_stream = [torch.cuda.Stream(device=out.device) for _ in peers[1]]
                sides = self._copy_stream
                npiece = max(1, min(self.pieces, (c1 - c0) // (64 * self.align)))
                units = (c1 - c0 + self.align - 1) // self.align
                for i in range(npiece):
                    a = c0 + min(c1 - c0, (units * i // npiece) * self.align)
                    b = c0 + min(c1 - c0, (units * (i + 1) // npiece) * self.align)
                    if b <= a:
                        continue
                    self.engine.apply_weights_cols(x, w, a, b, out)
                    ev = torch.cuda.Event()
                    ev.record(main)
                    for q in range(len(sides)):               # start with a different peer on every rank: no incast on one GPU
                        j = (q + self.rank) % len(sides)
                        side = sides[j]
                        side.wait_event(ev)
                        if hasattr(self.engine, "peer_copy_cols"):
                            self.engine.peer_copy_cols(peers[1][j], out, a, b, stream=side)     # copy engine, strided
                        else:
                            with torch.cuda.stream(side):
                                pflat[j][:, a:b].copy_(flat[:, a:b], non_blocking=True)
                for side in sides:
                    main.wait_stream(side)
            peers[2].barrier()                                # every rank's stores have landed before anyone reads its output
            return out
        self.engine.apply_weights_cols(x, w, c0, c1, out)
        if gather and self.world > 1:
            self._gather(out, cols)
        return out

    def _gather(self, out, cols):
        """All-gather of the analysed column ranges.  The (rows, N) analysis is row-major, so a rank's columns are ``rows``
        strided pieces: with equal range widths (all boundaries multiples of ``align``, N divisible) the ranges are exchanged
        by ONE grouped collective of per-row-block all-gathers straight into ``out`` — no staging buffer, no copies; otherwise
        through a rank-major staging buffer."""
        rows = out.shape[0] * out.shape[1]
        n = out.shape[-1]
        flat = out.view(rows, n)
        c0, c1 = cols[self.rank]
        widths = [b - a for a, b in cols]
        if self._recv is None or self._recv.shape != (self.world, rows, max(widths)) or self._recv.dtype != out.dtype \
                or self._recv.device != out.device:
            self._recv = torch.empty((self.world, rows, max(widths)), dtype=out.dtype, device=out.device)
        mine = self._recv[self.rank]
        mine[:, :c1 - c0].copy_(flat[:, c0:c1])
        dist.all_gather_into_tensor(self._recv.view(self.world * rows, max(widths)), mine, group=self.group)
        for r, (a, b) in enumerate(cols):
            if r != self.rank and b > a:
                flat[:, a:b].copy_(self._recv[r, :, :b - a])
        return out
