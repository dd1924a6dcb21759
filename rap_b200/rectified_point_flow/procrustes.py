"""`rectified_point_flow.procrustes` on the fused sm_100a Procrustes kernel (reference procrustes.py:6-118).

Same functions, arguments and return shapes as the reference.  Differences are only in how the work is
done: one kernel launch for all parts and no host synchronisation instead of a Python loop with a
cuSOLVER SVD + det() + device sync per (scene, part) slot.
"""
from __future__ import annotations

import torch

from .._lib import check, lib, ptr, stream_ptr


def _prep(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("rap_b200.procrustes runs on CUDA tensors only (no CPU fallback)")
    return t.detach().to(torch.float32).reshape(-1, 3).contiguous()


def _slot_offsets(points_per_part: torch.Tensor, device, n_points: int | None = None) -> torch.Tensor:
    """cu_seqlens over ALL (b,p) slots in packing order; empty slots become zero-length parts that the kernel
    skips, so no host sync (`ppp[ppp>0]`) is needed and slot index == b*P+p.

    The reference asserts `sum(points_per_part) == N` (utils/point_clouds.py:41-44).  A host-resident table is checked
    here the same way; a device-resident one is not read back (that would be the host sync this path removes) -- the
    kernel gets N instead, skips any part that reaches past it and flags it in the device error word."""
    if n_points is not None and not points_per_part.is_cuda and int(points_per_part.sum()) != n_points:
        raise ValueError(f"points_per_part sums to {int(points_per_part.sum())}, but the point cloud has {n_points} points")
    cu = torch.zeros(points_per_part.numel() + 1, dtype=torch.int32, device=device)
    cu[1:] = torch.cumsum(points_per_part.to(device).reshape(-1), 0)
    return cu


def solve_procrustes(source_pcd: torch.Tensor, target_pcd: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """min_{R,t} |R source + t - target|^2 for one (N,3) pair -> R (3,3), t (3,)  (reference procrustes.py:6-37)."""
    s, g = _prep(source_pcd), _prep(target_pcd)
    if s.shape != g.shape:
        raise ValueError("source and target must have the same shape")
    cu = torch.tensor([0, s.shape[0]], dtype=torch.int32, device=s.device)
    R = torch.zeros(1, 3, 3, device=s.device)
    t = torch.zeros(1, 3, device=s.device)
    check(lib.rap_procrustes_fit(ptr(s), ptr(g), ptr(cu), None, 1, s.shape[0], ptr(R), ptr(t), stream_ptr()), "rap_procrustes_fit")
    return R[0], t[0]


def fit_transformations(source_pcds: torch.Tensor, target_pcds: torch.Tensor, points_per_part: torch.Tensor,
                        cu_seqlens_batch: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor]:
    """Per-part rigid fit -> rotations (B,P,3,3), translations (B,P,3); zeros for empty slots
    (reference procrustes.py:40-84).  Accepts packed (TP,3) or (B,N,3) clouds."""
    if source_pcds.ndim == 2 and cu_seqlens_batch is None:
        raise ValueError("cu_seqlens_batch is required when pointclouds has shape (TP, 3)")
    s, g = _prep(source_pcds), _prep(target_pcds)
    B, P = points_per_part.shape
    cu = _slot_offsets(points_per_part, s.device, s.shape[0])
    R = torch.zeros(B, P, 3, 3, device=s.device)
    t = torch.zeros(B, P, 3, device=s.device)
    check(lib.rap_procrustes_fit(ptr(s), ptr(g), ptr(cu), None, B * P, s.shape[0], ptr(R), ptr(t), stream_ptr()), "rap_procrustes_fit")
    return R, t


def rigidify_prediction_with_procrustes(prediction: torch.Tensor, condition: torch.Tensor, points_per_part: torch.Tensor,
                                        cu_seqlens_batch: torch.Tensor | None = None) -> torch.Tensor:
    """out[part] = condition_part @ R^T + t with (R,t) fitted condition_part -> prediction_part
    (reference procrustes.py:86-118).  Same shape as `prediction`."""
    if prediction.ndim == 2 and cu_seqlens_batch is None:
        raise ValueError("cu_seqlens_batch is required when pointclouds has shape (TP, 3)")
    p, c = _prep(prediction), _prep(condition)
    B, P = points_per_part.shape
    cu = _slot_offsets(points_per_part, p.device, p.shape[0])
    out = torch.zeros_like(p)
    check(lib.rap_procrustes_rigidify(ptr(p), ptr(c), ptr(cu), B * P, p.shape[0], ptr(out), None, 1.0, 0.0, stream_ptr()), "rap_procrustes_rigidify")
    return out.view(prediction.shape)


def rigidify_and_interpolate(prediction, condition, x_1, points_per_part, w_rigid: float, w_noise: float) -> torch.Tensor:
    """Fused sampler.py:59-60: rigid(prediction) * w_rigid + x_1 * w_noise in the same launch."""
    p, c, n = _prep(prediction), _prep(condition), _prep(x_1)
    B, P = points_per_part.shape
    cu = _slot_offsets(points_per_part, p.device, p.shape[0])
    out = torch.empty_like(p)
    check(lib.rap_procrustes_rigidify(ptr(p), ptr(c), ptr(cu), B * P, p.shape[0], ptr(out), ptr(n), float(w_rigid), float(w_noise), stream_ptr()),
          "rap_procrustes_rigidify")
    return out.view(prediction.shape)
