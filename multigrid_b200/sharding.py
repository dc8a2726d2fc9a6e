"""Env-axis sharding across the GPUs of one box.

The batch is embarrassingly parallel: rank r owns the contiguous global env ids
[r*E/G, (r+1)*E/G); per-env seeds are a function of the GLOBAL id, so the results of env e do
not depend on how many GPUs the batch is split over. There is NO collective on the step path
(no NCCL traffic per step); torch.distributed is used only to aggregate counters / timings.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_bounds(num_envs_total: int, world_size: int, rank: int) -> tuple[int, int]:
    """[first, last) global env ids of `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(num_envs_total), int(world_size))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def make_sharded(env_id: str, num_envs_total: int, agents: int = 1, **kwargs):
    """This rank's shard of a global batch of `num_envs_total` envs (`envs.make` underneath)."""
    from .envs import make
    rank, world_size, local_rank = world()
    first, last = shard_bounds(num_envs_total, world_size, rank)
    kwargs.setdefault("device", f"cuda:{local_rank}")
    return make(env_id, agents=agents, num_envs=last - first, first_env=first, **kwargs)


def max_over_ranks(values, device) -> list[float]:
    """Element-wise max over ranks (timings are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def sum_over_ranks(values, device) -> list[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]
