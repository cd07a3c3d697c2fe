"""Multi-GPU plumbing: independent worlds are partitioned contiguously over ranks (one
process per GPU); there is no per-step inter-GPU traffic.  The only collective is the
end-of-run reduce of {checksum, energy, counters, max time} over torch.distributed (NCCL
over NVLink on GPUs; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple


def shard_range(n_worlds: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous shard [first, first+count) of rank: GPU g of G owns worlds
    [g*W/G, (g+1)*W/G) with the remainder spread over the first ranks."""
    base, rem = divmod(n_worlds, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def reduce_run(checksum: int, energy: float, counters: Dict[str, int], elapsed_ms: float, device) -> dict:
    """All-reduce the end-of-run quantities.  The u64 checksum is summed mod 2^64 (carried as
    two 32-bit halves in int64 so that the sum can not overflow); energy in float64; time MAX."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return {"checksum": checksum % (1 << 64), "energy": energy, "counters": dict(counters), "max_ms": elapsed_ms}
    keys = sorted(counters)
    ints = torch.tensor([checksum & 0xFFFFFFFF, (checksum >> 32) & 0xFFFFFFFF] + [int(counters[k]) for k in keys], dtype=torch.int64, device=device)
    dist.all_reduce(ints, op=dist.ReduceOp.SUM)
    fl = torch.tensor([energy], dtype=torch.float64, device=device)
    dist.all_reduce(fl, op=dist.ReduceOp.SUM)
    tm = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    v = ints.tolist()
    total = (v[0] + (v[1] << 32)) % (1 << 64)
    return {"checksum": total, "energy": float(fl.item()), "counters": {k: v[2 + i] for i, k in enumerate(keys)}, "max_ms": float(tm.item())}
