"""Instance sharding across GPUs: controller instances are independent, so a batch splits into
contiguous per-rank blocks and the solve needs NO data-path collective.  torch.distributed (NCCL on the GPU
box, gloo in the CPU tests) carries only the timing / statistics reduction below."""
from __future__ import annotations


def shard_range(nbatch: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the instances rank `rank` owns: contiguous blocks of ceil(nbatch / world), last ranks may be short/empty."""
    if world < 1 or not (0 <= rank < world) or nbatch < 0:
        raise ValueError("bad shard request")
    per = -(-nbatch // world)
    lo = min(rank * per, nbatch)
    return lo, min(lo + per, nbatch)


def reduce_stats(elapsed_ms: float, counts, device=None):
    """(max over ranks of elapsed_ms, element-wise sum over ranks of counts).  Works without an initialised
    process group (single process)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    c = torch.tensor([float(v) for v in counts], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), [float(v) for v in c.tolist()]
