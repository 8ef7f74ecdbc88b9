"""Data-parallel sharding of sub-scenes over GPUs.

The reference is single-process / single-GPU (`sim_backend="physx_cuda:n"`, mani_skill/envs/utils/system/backend.py:46-68).
Sub-scenes are independent, so rank r of G owns the contiguous env block [r*N/G, (r+1)*N/G) with its own world, buffers,
cameras and stream; physics and rendering need no communication.  The only collective is the optional all-gather of the
flattened observation when the caller wants the global batch on every rank.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of global env ids owned by `rank` (sizes differ by at most one)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(total_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seeds(base_seed: int, total_envs: int, rank: int, world_size: int):
    """Per-env seeds keep the GLOBAL env id (sapien_env.py:321 seeds 2022+i), so an env's seed does not depend on G."""
    lo, hi = shard_range(total_envs, rank, world_size)
    return [base_seed + i for i in range(lo, hi)]


class ObsGather:
    """all_gather_into_tensor of a [n_local, D] observation into a persistent [n_total, D] buffer (equal shards)."""

    def __init__(self, n_local: int, dim: int, dtype=torch.float32, device=None, group: Optional[dist.ProcessGroup] = None):
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buf = torch.empty((self.world_size * n_local, dim), dtype=dtype, device=device)

    def __call__(self, obs: torch.Tensor) -> torch.Tensor:
        if self.world_size == 1:
            self.buf.copy_(obs)
            return self.buf
        dist.all_gather_into_tensor(self.buf, obs.contiguous(), group=self.group)
        return self.buf
