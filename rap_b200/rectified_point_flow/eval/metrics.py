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
    check(lib.rap_rigid_sse(ptr(s), ptr(g), ptr(cu), None, B * P, s.shape[0], ptr(R), ptr(t), ptr(sse), stream_ptr()), "rap_rigid_sse")
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
    check(lib.rap_procrustes_fit_sse(ptr(s), ptr(g), ptr(cu), None, B * P, s.shape[0], ptr(R), ptr(t), ptr(sse), stream_ptr()), "rap_procrustes_fit_sse")
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
        check(lib.rap_procrustes_fit_sse(ptr(s), ptr(traj[i]), ptr(cu), None, B * P, s.shape[0], ptr(R), ptr(t), ptr(sse[i]), stream_ptr()),
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


# ------------------------------------------------------------------------------------------------ metrics against ground truth (rank 4)
def _segment_nn(query: torch.Tensor, target: torch.Tensor, cu: torch.Tensor, labels: torch.Tensor | None = None) -> torch.Tensor:
    q, t = _prep(query), _prep(target)
    cu32 = cu.detach().to(q.device, torch.int32).contiguous()
    lab = labels.detach().to(q.device, torch.int32).contiguous() if labels is not None else None
    out = torch.empty(q.shape[0], device=q.device)
    check(lib.rap_segment_nn(ptr(q), ptr(t), ptr(cu32), cu32.numel() - 1, ptr(lab), ptr(lab), ptr(out), stream_ptr()), "rap_segment_nn")
    return out


def _segment_mean(values: torch.Tensor, cu: torch.Tensor) -> torch.Tensor:
    cu = cu.to(values.device, torch.long)
    n = (cu[1:] - cu[:-1])
    seg = torch.repeat_interleave(torch.arange(n.numel(), device=values.device), n)
    return torch.zeros(n.numel(), device=values.device, dtype=values.dtype).index_add_(0, seg, values) / n.clamp(min=1)


def compute_cd(pointclouds_gt: torch.Tensor, pointclouds_pred: torch.Tensor, cu_seqlens_batch: torch.Tensor,
               anchor_indices: torch.Tensor | None = None) -> torch.Tensor:
    """(B,) whole-object chamfer RMSE, sqrt(0.5 * (mean_x min_y |x-y|^2 + mean_y min_x |x-y|^2)) (reference eval/metrics.py:14-48,
    pytorch3d chamfer_distance with norm=2, point_reduction='mean'); both directions on the nearest-neighbour kernel."""
    g, p = pointclouds_gt.reshape(-1, 3), pointclouds_pred.reshape(-1, 3)
    cd = _segment_mean(_segment_nn(g, p, cu_seqlens_batch), cu_seqlens_batch) + _segment_mean(_segment_nn(p, g, cu_seqlens_batch), cu_seqlens_batch)
    return (0.5 * cd).sqrt()


def compute_overlap_ratio(pointclouds_pred: torch.Tensor, points_per_part: torch.Tensor, cu_seqlens_batch: torch.Tensor | None = None,
                          taus=(0.005, 0.01, 0.02)) -> torch.Tensor:
    """(T,B) share of points with a point of a DIFFERENT part within tau (reference eval/metrics.py:625-691); 0 for objects with fewer
    than two parts or points."""
    pred = _prep(pointclouds_pred)
    B, P = points_per_part.shape
    ppp = points_per_part.to(pred.device, torch.long)
    cu = cu_seqlens_batch.to(pred.device, torch.long) if cu_seqlens_batch is not None else torch.cat([ppp.new_zeros(1), ppp.sum(1).cumsum(0)])
    part_of_point = torch.repeat_interleave(torch.arange(P, device=pred.device).repeat(B), ppp.reshape(-1))
    d2 = _segment_nn(pred, pred, cu, part_of_point)
    tau_list = [float(taus)] if isinstance(taus, (float, int)) else [float(t) for t in taus]
    out = torch.stack([_segment_mean((d2.sqrt() <= tau).to(pred.dtype), cu) for tau in tau_list])
    usable = ((cu[1:] - cu[:-1]) > 1) & ((ppp > 0).sum(1) > 1)
    return out * usable


def _rot_angle_deg(delta_R: torch.Tensor) -> torch.Tensor:
    cos = (0.5 * (delta_R.diagonal(dim1=-2, dim2=-1).sum(-1) - 1)).clamp(-1.0, 1.0)
    return torch.rad2deg(torch.acos(cos))


def compute_transform_errors_direct(rotations_gt, translations_gt, rotations_pred, translations_pred, points_per_part, matched_part_ids=None,
                                    scale=None):
    """Mean per-object rotation (deg) / translation (m) error over all non-empty parts (reference eval/metrics.py:305-384), batched."""
    B, P = points_per_part.shape
    dev = rotations_gt.device
    if matched_part_ids is not None:
        bi = torch.arange(B, device=dev)[:, None]
        rotations_pred, translations_pred = rotations_pred[bi, matched_part_ids], translations_pred[bi, matched_part_ids]
    scale = torch.ones(B, device=dev) if scale is None else scale.to(dev)
    valid = points_per_part.to(dev) != 0
    rot = _rot_angle_deg(rotations_gt.float().transpose(-1, -2) @ rotations_pred.float()) * valid
    tr = ((translations_pred.float() - translations_gt.float()) * scale.view(-1, 1, 1)).norm(dim=-1) * valid
    n = valid.sum(1)
    return rot.sum(1) / n, tr.sum(1) / n


def compute_transform_errors(pointclouds, pointclouds_gt, rotations_gt, translations_gt, rotations_pred, translations_pred, points_per_part,
                             anchor_part, matched_part_ids=None, scale=None, cu_seqlens_batch=None, use_icp: bool = False):
    """Mean per-object rotation / translation error of the non-anchor parts, in the frame of the (first) anchor part (reference
    eval/metrics.py:165-303, use_icp=False branch -- the one the registration task uses), batched over (B,P)."""
    if use_icp:
        raise NotImplementedError("use_icp=True (interchangeable parts) is not used by the registration task")
    B, P = points_per_part.shape
    dev = rotations_gt.device
    if matched_part_ids is not None:
        bi = torch.arange(B, device=dev)[:, None]
        rotations_pred, translations_pred = rotations_pred[bi, matched_part_ids], translations_pred[bi, matched_part_ids]
    scale = torch.ones(B, device=dev) if scale is None else scale.to(dev)
    anchor = anchor_part.to(dev).bool()
    has = anchor.any(1)
    first = torch.argmax(anchor.int(), dim=1)
    bi = torch.arange(B, device=dev)
    eye = torch.eye(3, device=dev).expand(B, 3, 3)

    def anchor_inverse(R, t):
        Ra, ta = R[bi, first].float(), t[bi, first].float()
        Ri = torch.where(has[:, None, None], Ra.transpose(-1, -2), eye)
        ti = torch.where(has[:, None], -(Ri @ ta[..., None])[..., 0], torch.zeros(B, 3, device=dev))
        return Ri[:, None], ti[:, None]

    Rg_i, tg_i = anchor_inverse(rotations_gt, translations_gt)
    Rp_i, tp_i = anchor_inverse(rotations_pred, translations_pred)
    R_gt_rel, t_gt_rel = Rg_i @ rotations_gt.float(), (Rg_i @ translations_gt.float()[..., None])[..., 0] + tg_i
    R_pr_rel, t_pr_rel = Rp_i @ rotations_pred.float(), (Rp_i @ translations_pred.float()[..., None])[..., 0] + tp_i
    valid = (points_per_part.to(dev) != 0) & ~anchor
    rot = _rot_angle_deg(R_gt_rel.transpose(-1, -2) @ R_pr_rel) * valid
    tr = ((t_pr_rel - t_gt_rel) * scale.view(-1, 1, 1)).norm(dim=-1) * valid
    n = valid.sum(1)
    return rot.sum(1) / n, tr.sum(1) / n
