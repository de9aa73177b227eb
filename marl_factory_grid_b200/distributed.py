"""Multi-GPU plumbing: envs shard trivially (no per-step collective); only episode statistics are reduced.

One process per GPU (torchrun).  Rank r owns the contiguous global env ids [r * n_local, (r + 1) * n_local); the
Philox streams are keyed by the GLOBAL env id, so a sharded run reproduces the single-process run env for env.
The only exchange of the path is one small all-reduce (SUM) of the statistics vector over NCCL / NVLink (gloo on CPU).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .abi import ST_RETURN_SUM

N_COUNTERS = ST_RETURN_SUM          # integer counters come first in the statistics vector; f64 sums follow


def shard(rank: int, world: int, n_total: int) -> Tuple[int, int]:
    """(env_id_offset, n_local) of a contiguous block partition; the first `n_total % world` ranks get one more env."""
    base, rem = divmod(int(n_total), int(world))
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, n_local


def allreduce_stats(stats: np.ndarray, device=None) -> np.ndarray:
    """Sum the engine's statistics vector over all ranks (int64 counters and the f64 return sums separately)."""
    import torch
    import torch.distributed as dist
    stats = np.asarray(stats, np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return stats.copy()
    counters = torch.as_tensor(stats[:N_COUNTERS].copy(), device=device)
    sums = torch.as_tensor(stats[N_COUNTERS:].view(np.float64).copy(), device=device)
    dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    out = stats.copy()
    out[:N_COUNTERS] = counters.cpu().numpy()
    out[N_COUNTERS:] = sums.cpu().numpy().view(np.int64)
    return out


def allreduce_max(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
