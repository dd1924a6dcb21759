"""`rectified_point_flow.eval.metrics.compute_rigidity_rmse` on the fused Procrustes kernel (reference eval/metrics.py:511-622) --
the metric `test_step` uses to pick the best of n generations (modeling.py:456-547).  Only this function of the reference's metrics
module is on the path; chamfer / pose-error metrics need ground truth and stay with the reference."""
from __future__ import annotations

import torch

from ..._lib import check, lib, ptr, stream_ptr
from ..procrustes import _prep, _slot_offsets


def compute_rigidity_rmse(pointclouds_input: torch.Tensor, pointclouds_pred: torch.Tensor, rotations_pred: torch.Tensor,
                          translations_pred: torch.Tensor, points_per_part: torch.Tensor, cu_seqlens_batch: torch.Tensor | None = None,
                          scales: torch.Tensor | None = None, average_per_part: bool = False) -> torch.Tensor:
    """(B,) RMSE between `input @ R^T + t` and the generated cloud, per object; point-weighted by default, or the mean of per-part
    RMSEs; inf for objects without parts; multiplied by `scales` if given.  One launch, no host loop."""
    s, g = _prep(pointclouds_input), _prep(pointclouds_pred)
    B, P = points_per_part.shape
    cu = _slot_offsets(points_per_part, s.device)
    R = rotations_pred.detach().to(s.device, torch.float32).contiguous()
    t = translations_pred.detach().to(s.device, torch.float32).contiguous()
    sse = torch.zeros(B, P, device=s.device)
    check(lib.rap_rigid_sse(ptr(s), ptr(g), ptr(cu), None, B * P, ptr(R), ptr(t), ptr(sse), stream_ptr()), "rap_rigid_sse")
    n = points_per_part.to(s.device, torch.float32)
    valid = n > 0
    if average_per_part:
        part_rmse = torch.sqrt(sse / n.clamp(min=1))
        rmse = (part_rmse * valid).sum(1) / valid.sum(1).clamp(min=1)
    else:
        rmse = torch.sqrt(sse.sum(1) / n.sum(1).clamp(min=1))
    rmse = torch.where(valid.any(1), rmse, torch.full_like(rmse, float("inf")))
    return rmse * scales.to(rmse.device) if scales is not None else rmse


def fit_and_rigidity_sse(source: torch.Tensor, target: torch.Tensor, points_per_part: torch.Tensor):
    """Fused `fit_transformations` + squared rigid residual per part (what modeling.py:473-487 does per generation and step in two
    host loops): returns R (B,P,3,3), t (B,P,3), sse (B,P)."""
    s, g = _prep(source), _prep(target)
    B, P = points_per_part.shape
    cu = _slot_offsets(points_per_part, s.device)
    R = torch.zeros(B, P, 3, 3, device=s.device)
    t = torch.zeros(B, P, 3, device=s.device)
    sse = torch.zeros(B, P, device=s.device)
    check(lib.rap_procrustes_fit_sse(ptr(s), ptr(g), ptr(cu), None, B * P, ptr(R), ptr(t), ptr(sse), stream_ptr()), "rap_procrustes_fit_sse")
    return R, t, sse


def trajectory_rigidity_rmse(condition: torch.Tensor, trajectory: torch.Tensor, points_per_part: torch.Tensor,
                             scales: torch.Tensor | None = None) -> torch.Tensor:
    """Mean over the steps of an end-point trajectory (S,T,3) of the per-object rigidity RMSE, each step with its own fit -- the
    `use_average_rigidity_rmse` selection score of modeling.py:462-500, one fused launch per step (no (B,P) host loops)."""
    s = _prep(condition)
    B, P = points_per_part.shape
    cu = _slot_offsets(points_per_part, s.device)
    S = trajectory.shape[0]
    traj = trajectory.detach().to(s.device, torch.float32).reshape(S, -1, 3).contiguous()
    R = torch.zeros(B, P, 3, 3, device=s.device)
    t = torch.zeros(B, P, 3, device=s.device)
    sse = torch.zeros(S, B, P, device=s.device)
    for i in range(S):
        check(lib.rap_procrustes_fit_sse(ptr(s), ptr(traj[i]), ptr(cu), None, B * P, ptr(R), ptr(t), ptr(sse[i]), stream_ptr()),
              "rap_procrustes_fit_sse")
    n = points_per_part.to(s.device, torch.float32).sum(1)
    rmse = torch.sqrt(sse.sum(2) / n.clamp(min=1))                       # (S,B)
    rmse = torch.where(n > 0, rmse, torch.full_like(rmse, float("inf")))
    if scales is not None:
        rmse = rmse * scales.to(rmse.device)
    return rmse.mean(0) if S else torch.full((B,), float("inf"), device=s.device)


def select_generation(scores: torch.Tensor) -> torch.Tensor:
    """(G,B) rigidity scores -> (B,) index of the generation to keep (argmin, modeling.py:518)."""
    return torch.argmin(scores, dim=0)
