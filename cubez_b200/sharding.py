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


def pin_to_gpu_numa_node(local_rank: int) -> dict:
    """Bind this rank's host threads to the CPUs of the NUMA node its GPU hangs off (sysfs; no libnuma needed), BEFORE
    any pinned buffer is allocated: page-locked memory is then first-touched on that node, so the H2D / D2H copies of the
    end-to-end arms do not cross the socket interconnect.  Returns what was found (reported in bench.py's config)."""
    import os
    import subprocess
    info = {"gpu": local_rank, "node": None, "cpus": None, "pinned": False}
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = "0000:" + bus.split(":", 1)[1]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        info["node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpulist = f.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        info["cpus"] = cpulist
        n_nodes = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])
        info["nodes"] = n_nodes
        if allowed and n_nodes > 1:
            os.sched_setaffinity(0, allowed)
            info["pinned"] = True
    except Exception as e:      # sysfs not readable in a container: leave the affinity alone
        info["error"] = type(e).__name__
    return info
