#!/usr/bin/env python
"""Transport of the per-analysis inputs from rank 0 to all ranks (cfg3: Yn 1.0 GB, state 0.4 GB, d 20 MB, obs coordinates
40 MB): four NCCL broadcasts, one flat broadcast, scatter + in-place all-gather.  torchrun, one process per GPU."""
import json
import os
import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    sizes = [50 * 2_500_000, 50 * 1_000_000, 2_500_000, 2 * 2_500_000]          # doubles
    tensors = [torch.zeros(n, dtype=torch.float64, device=dev) for n in sizes]
    total = sum(sizes)
    per = (total + world - 1) // world
    flat = torch.zeros(per * world, dtype=torch.float64, device=dev)
    out = {}

    def timed(name, fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = float(t)

    timed("four_broadcasts", lambda: [dist.broadcast(t, 0) for t in tensors])
    timed("one_flat_broadcast", lambda: dist.broadcast(flat, 0))

    def scatter_allgather():
        mine = flat[rank * per:(rank + 1) * per]
        dist.scatter(mine, [flat[r * per:(r + 1) * per] for r in range(world)] if rank == 0 else None, src=0)
        dist.all_gather_into_tensor(flat, mine)
    timed("scatter_then_allgather", scatter_allgather)
    timed("allgather_only", lambda: dist.all_gather_into_tensor(flat, flat[rank * per:(rank + 1) * per]))
    if rank == 0:
        out["bytes"] = total * 8
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
