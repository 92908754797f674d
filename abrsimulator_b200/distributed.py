"""Multi-GPU sharding: sessions are independent, so ranks own contiguous blocks and never exchange
data while stepping; the only collective is the final QoE/statistics reduction (SPEC.md §6).

The per-rank statistic vectors are all-gathered and summed in rank order on every rank, so the
result is bit-reproducible run to run (an fp64 SUM inside NCCL's ring/tree has no fixed order).
Works with NCCL (one process per GPU over NVLink) and gloo (CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block [lo, hi) of rank; the first n_total % world ranks own one extra session."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sum a per-rank statistics vector over all ranks, deterministically (rank order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats.clone()
    world = dist.get_world_size(group)
    parts = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(parts, stats.contiguous(), group=group)
    total = parts[0].clone()
    for p in parts[1:]:
        total += p
    return total


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Max of a scalar (e.g. elapsed milliseconds) over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def bind_to_gpu_cpus(device_index: int) -> list | None:
    """Restrict this process to the CPU cores next to GPU ``device_index`` (NVML's ideal CPU affinity: the cores of
    the NUMA node its PCIe root hangs off).  One process per GPU: host buffers allocated afterwards are first-touched
    on that node, so the kernels' zero-copy PCIe reads of pinned buffers and the launch path stay local.  Returns the
    core list, or None when NVML or the affinity call is unavailable or fewer than four of those cores are usable by
    this process (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = device_index
        if visible:                                    # NVML ignores CUDA_VISIBLE_DEVICES
            ids = [x.strip() for x in visible.split(",") if x.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                index = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cores = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if len(allowed) < 4:                           # leave room for the driver's and NCCL's helper threads
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
