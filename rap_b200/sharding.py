"""Scene sharding across ranks: scenes are independent (part attention stays inside a view, global attention and
AdaLN inside a scene, Procrustes inside a view), so rank r registers its own block of scenes with replicated weights, there is no
data-path collective, and the only exchange is ONE all-gather of the recovered poses per batch -- what the reference gets from
Lightning's DistributedSampler index sharding (data/datamodule.py:103-106) plus its evaluator's gather.

`Registrar.register_sharded` (engine.py) is the product entry point; this module holds the host logic it is built from, which the
gloo world-size-2 CPU tests exercise without a GPU.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.distributed as dist


def shard_range(n_scenes: int, rank: int, world: int) -> range:
    """Contiguous block of scene indices for `rank`; remainders go to the lowest ranks (all scenes covered once)."""
    base, rem = divmod(n_scenes, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_counts(n_scenes: int, world: int) -> list[int]:
    """Scenes per rank under `shard_range` -- every rank can compute everybody's count, so uneven shards need no size exchange."""
    return [len(shard_range(n_scenes, r, world)) for r in range(world)]


def slice_scenes(batch, s0: int, s1: int):
    """Scenes [s0, s1) of a packed batch (attributes cond, features, scales, anchor_indices, points_per_part, cu_seqlens_batch, x_1 --
    the layout `variable_collate_fn` + `_prepare_data` produce, reference data/datamodule.py:169-198, modeling.py:203-231) as a new
    packed batch with re-based offsets.  Works on host or device tensors; tensor slices are views."""
    cu = batch.cu_seqlens_batch
    a, b = int(cu[s0]), int(cu[s1])
    ppp = batch.points_per_part[s0:s1]
    valid = ppp[ppp > 0]
    cu_p = torch.zeros(valid.numel() + 1, dtype=torch.int32, device=valid.device)
    cu_p[1:] = torch.cumsum(valid, 0)
    out = SimpleNamespace(cond=batch.cond[a:b], features=batch.features[a:b] if batch.features is not None else None, scales=batch.scales[s0:s1],
                          anchor_indices=batch.anchor_indices[a:b], points_per_part=ppp, cu_seqlens_batch=(cu[s0:s1 + 1] - cu[s0]).to(torch.int32),
                          cu_seqlens_part=cu_p, x_1=batch.x_1[a:b])
    if getattr(batch, "gt", None) is not None:
        out.gt = batch.gt[a:b]
    return out


def micro_batches(cu_seqlens_batch, s0: int, s1: int, max_points: int | None) -> list[tuple[int, int]]:
    """Cuts scenes [s0, s1) into consecutive micro-batches of at most `max_points` packed points (the reference bounds a batch the same
    way with `max_points_per_batch`, config/RAP_inference.yaml:36); a single scene above the budget gets a micro-batch of its own."""
    if s1 <= s0:
        return []
    if not max_points:
        return [(s0, s1)]
    cu = [int(v) for v in cu_seqlens_batch[s0:s1 + 1]]
    out, start = [], 0
    for i in range(1, len(cu)):
        if cu[i] - cu[start] > max_points and i - 1 > start:
            out.append((s0 + start, s0 + i - 1))
            start = i - 1
    out.append((s0 + start, s1))
    return out


def gather_poses(rotations: torch.Tensor, translations: torch.Tensor, group=None, counts: list[int] | None = None) -> torch.Tensor:
    """(B_local,P,3,3),(B_local,P,3) on every rank -> (sum B_local, P, 12) [R row-major | t] on every rank, rank order.
    NCCL on GPUs (one all-gather, a few KB: latency-bound, NVLink bandwidth irrelevant); gloo in CPU tests.
    `counts` = scenes per rank when shards are uneven (`shard_counts`): the local block is zero-padded to the largest shard for the
    collective and the padding rows are dropped from the result.  P must be the same on every rank (it is a property of the global
    batch's `points_per_part`)."""
    local = torch.cat([rotations.flatten(2), translations], dim=-1).contiguous()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    if counts is None:
        counts = [local.shape[0]] * world
    if len(counts) != world or counts[dist.get_rank(group)] != local.shape[0]:
        raise ValueError(f"gather_poses: counts {counts} do not match world size {world} / the local shard of {local.shape[0]} scenes")
    width = max(counts)
    if width == 0:
        return local
    if local.shape[0] < width:
        local = torch.cat([local, local.new_zeros(width - local.shape[0], *local.shape[1:])])
    out = torch.empty(world * width, *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    if all(c == width for c in counts):
        return out
    return torch.cat([out[r * width:r * width + c] for r, c in enumerate(counts)])
