"""Host-side sharding rules of the multi-GPU configurations.  No device code here; the same
functions run under NCCL on the GPU box and under gloo in the CPU tests.

  * C4 replica ensembles: replicas are independent units, rank g takes the contiguous block
    [g*R/G, (g+1)*R/G); replica r integrates with seed base_seed + r wherever it runs, so a sharded ensemble is
    the same ensemble.  No collective on the data path.
  * C2-style replicas (bench.py --gpus N): one independent system per rank, seed base_seed + rank.
  * timing: device seconds reduced with MAX over ranks; throughput = units of all ranks / that time.
"""
from __future__ import annotations


def replica_shard(total: int, world: int, rank: int):
    """(first replica, count) of rank `rank`: contiguous, sizes differ by at most one, every replica exactly once"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def replica_seeds(base_seed: int, total: int, world: int, rank: int):
    first, count = replica_shard(total, world, rank)
    return [base_seed + r for r in range(first, first + count)]


def max_over_ranks(values, device=None):
    """element-wise MAX of a list of floats over all ranks of the default process group (identity without one)"""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_concat(array, device=None):
    """concatenation over ranks of per-rank 1-D float arrays of possibly different length (rank order)"""
    import numpy as np
    import torch
    import torch.distributed as dist

    a = np.asarray(array, dtype=np.float64).reshape(-1)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return a
    world = dist.get_world_size()
    n = torch.tensor([a.size], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(s.item() for s in sizes))
    buf = torch.zeros(m, dtype=torch.float64, device=device)
    buf[: a.size] = torch.from_numpy(a).to(buf.device)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return np.concatenate([p[: int(s.item())].cpu().numpy() for p, s in zip(parts, sizes)])
