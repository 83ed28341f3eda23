"""Multi-GPU partitioning of pair batches.

Pair batches are independent units: each rank (one process per GPU) aligns a contiguous range of
pairs and nothing crosses GPUs on the data path -- the only collectives are the timing reduction
(max over ranks) and, optionally, gathering the small result records on rank 0.  Works with any
torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced-by-count range [lo, hi) of rank `rank` (uniform pair sizes)."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_cells(len_a: np.ndarray, len_b: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous ranges with (nearly) equal sum of m*n per rank -- ragged batches."""
    cells = np.asarray(len_a, dtype=np.float64) * np.asarray(len_b, dtype=np.float64)
    csum = np.concatenate([[0.0], np.cumsum(cells)])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(csum, total * r / world, side="left")))
    cuts.append(len(cells))
    cuts = np.maximum.accumulate(cuts)
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(world)]


def max_over_ranks(value: float, device=None) -> float:
    """Time-like quantities are reported as the max over ranks (the slowest GPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_items(items: np.ndarray, counts: List[int], device=None):
    """Gathers the per-pair result records (40 B each) of every rank on rank 0, in pair order.
    `counts[r]` = number of pairs of rank r.  Returns the concatenated array on rank 0, else None."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return items
    rank, world = dist.get_rank(), dist.get_world_size()
    width = max(max(counts), 1)
    buf = np.zeros((width, items.dtype.itemsize), dtype=np.uint8)
    if len(items):
        buf[:len(items)] = items.view(np.uint8).reshape(len(items), -1)
    mine = torch.from_numpy(buf).to(device or "cpu")
    outs = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, outs, dst=0)
    if rank != 0:
        return None
    parts = [outs[r].cpu().numpy()[:counts[r]].reshape(-1).view(items.dtype) for r in range(world) if counts[r] > 0]
    if not parts:
        return np.zeros(0, dtype=items.dtype)
    return np.concatenate(parts)
