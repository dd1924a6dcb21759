"""Packed-varlen helpers kept for callers of the reference API (reference utils/point_clouds.py:6-67,161-183).
Pure indexing glue; the CUDA kernels consume `cu_seqlens` directly and never call these."""
import torch


def split_parts(pointclouds: torch.Tensor, points_per_part: torch.Tensor, cu_seqlens_batch: torch.Tensor | None = None):
    """Packed (TP,3) [needs cu_seqlens_batch] or (B,N,3) cloud -> list over scenes of lists of (N_i,3) views
    (views of the input; empty part slots are skipped, as in the reference)."""
    B, _ = points_per_part.shape
    counts_per_scene = points_per_part.tolist()
    if pointclouds.ndim == 2 and pointclouds.shape[1] == 3:
        if cu_seqlens_batch is None:
            raise ValueError("cu_seqlens_batch is required when pointclouds has shape (TP, 3)")
        bounds = cu_seqlens_batch.tolist()
        scenes = [pointclouds[bounds[b]:bounds[b + 1]] for b in range(B)]
    else:
        scenes = list(pointclouds if pointclouds.ndim == 3 else pointclouds.view(B, -1, 3))
    parts = []
    for b, (scene, counts) in enumerate(zip(scenes, counts_per_scene)):
        assert sum(counts) == scene.size(0), (
            f"Mismatch detected: sum(counts)={sum(counts)} does not equal scene size {scene.size(0)} for batch {b}.")
        parts.append([s for s in torch.split(scene, counts, dim=0) if s.size(0) > 0])
    return parts


def repeat_by_cu_seqlens(x: torch.Tensor, cu_seqlens: torch.Tensor) -> torch.Tensor:
    """(B, ...) -> (TP, ...): row b repeated cu[b+1]-cu[b] times."""
    cu = cu_seqlens.to(device=x.device, dtype=torch.int32)
    assert cu.ndim == 1 and cu.numel() == x.shape[0] + 1, "Bad cu_seqlens shape"
    return x.repeat_interleave((cu[1:] - cu[:-1]).long(), dim=0)
