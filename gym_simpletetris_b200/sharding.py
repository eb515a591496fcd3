"""Multi-GPU: envs are independent, so they shard by contiguous global env-id ranges, one process per GPU,
with NO collective on the step path.  Piece streams are keyed by the GLOBAL env id, so env e plays the
same game for any number of GPUs.  The only collective is a SUM all-reduce of four episode counters."""
from __future__ import annotations

import os

import torch


def shard_bounds(num_envs: int, rank: int, world_size: int):
    """Global env ids [lo, hi) owned by `rank`: contiguous, sizes differ by at most one."""
    if not 0 <= rank < world_size:
        raise ValueError("rank outside [0, world_size)")
    base, rem = divmod(int(num_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_info():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    """SUM over ranks when torch.distributed is initialised (NCCL for CUDA tensors, gloo for CPU); identity
    otherwise.  Used for episode statistics only."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def make_sharded_vec_env(total_envs: int, rank=None, world_size=None, local_rank=None, **kwargs):
    """This rank's shard of a `total_envs`-wide VecEnv (global env ids keep their piece streams)."""
    from .vec_env import VecEnv

    r, w, lr = env_rank_info()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    local_rank = lr if local_rank is None else local_rank
    lo, hi = shard_bounds(total_envs, rank, world_size)
    kwargs.setdefault("device", f"cuda:{local_rank}")
    return VecEnv(hi - lo, env_id_base=kwargs.pop("env_id_base", 0) + lo, **kwargs)
