"""Scene sharding across ranks: scenes are independent, so rank r registers its own block of
scenes with replicated weights and the only collective is one all-gather of the recovered poses per batch."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_scenes: int, rank: int, world: int) -> range:
    """Contiguous block of scene indices for `rank`; remainders go to the lowest ranks (all scenes covered once)."""
    base, rem = divmod(n_scenes, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def gather_poses(rotations: torch.Tensor, translations: torch.Tensor, group=None) -> torch.Tensor:
    """(B_local,P,3,3),(B_local,P,3) on every rank -> (world*B_local, P, 12) [R row-major | t] on every rank.
    NCCL on GPUs (one all-gather, a few KB: latency-bound, NVLink bandwidth irrelevant); gloo in CPU tests."""
    local = torch.cat([rotations.flatten(2), translations], dim=-1).contiguous()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    out = torch.empty(world * local.shape[0], *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out
