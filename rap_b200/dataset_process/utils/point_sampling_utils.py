"""`dataset_process.utils.point_sampling_utils` (reference :11-309): occupied-voxel counts and batched FPS on the CUDA kernels, the
keypoint budget rules as host code with the reference's rounding and numpy-RNG call sequence."""
from __future__ import annotations

import ctypes as C
import logging
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from ..._lib import check, lib, ptr, stream_ptr

logger = logging.getLogger(__name__)


def calculate_voxel_coverage(points: np.ndarray, voxel_size: float) -> int:
    """Occupied voxels of floor(points / voxel_size), evaluated in the array's own precision (reference :11-31)."""
    if len(points) == 0:
        return 0
    pts = np.ascontiguousarray(points)
    if pts.dtype not in (np.float32, np.float64):
        pts = pts.astype(np.float64)
    t = torch.from_numpy(pts).cuda()
    count = C.c_int(0)
    check(lib.rap_voxel_coverage(ptr(t), t.shape[0], int(pts.dtype == np.float64), float(voxel_size), C.byref(count), stream_ptr()),
          "rap_voxel_coverage")
    return count.value


def calculate_adaptive_sample_count_per_part(parts_points: List[np.ndarray], voxel_size: float, voxel_ratio: float,
                                             min_points_per_part: int, max_sample_points: int) -> List[int]:
    """Keypoints per part = occupied voxels * ratio, clamped to [min_points_per_part, min(len, max_sample_points)] (reference :33-85)."""
    counts = []
    for pts in parts_points or []:
        if len(pts) == 0:
            counts.append(0)
            continue
        want = int(calculate_voxel_coverage(pts, voxel_size) * voxel_ratio)
        counts.append(min(max_sample_points, min(len(pts), max(min_points_per_part, want))))
    return counts


def _floors(pts_per_part: np.ndarray, num_points: int, min_points_per_part: int):
    """Per-part minimum (scaled down when the minima alone exceed the budget) and the budget left after it."""
    floor = np.minimum(min_points_per_part, pts_per_part)
    if floor.sum() > num_points:
        logger.warning("minimum points per part exceed the target; scaling down")
        floor = np.maximum(1, np.round(floor * (num_points / floor.sum())).astype(int))
    return floor, num_points - floor.sum()


def _allocate_by_point_count_logic(pts_per_part: np.ndarray, num_points: int, min_points_per_part: int) -> np.ndarray:
    """Reference :150-191: proportional to the points above the floor, then +-1 on uniformly drawn eligible parts."""
    floor, spare = _floors(pts_per_part, num_points, min_points_per_part)
    target = floor.copy()
    room = pts_per_part - floor
    if spare > 0 and room.sum() > 0:
        target = np.minimum(target + np.round(room * spare / room.sum()).astype(int), pts_per_part)
        diff = num_points - target.sum()
        while diff != 0:
            ok = target < pts_per_part if diff > 0 else target > floor
            if ok.sum() == 0:
                break
            step = 1 if diff > 0 else -1
            target[np.random.choice(np.where(ok)[0])] += step
            diff -= step
    return target


def _allocate_by_spatial_coverage_logic(coverage_per_part: np.ndarray, pts_per_part: np.ndarray, num_points: int,
                                        min_points_per_part: int) -> np.ndarray:
    """Reference :193-260: proportional to the occupied voxels, fix-up draws weighted by coverage (inverse when shrinking)."""
    floor, spare = _floors(pts_per_part, num_points, min_points_per_part)
    target = floor.copy()
    total = coverage_per_part.sum()
    if spare > 0 and total > 0:
        target += np.minimum(np.round(coverage_per_part / total * spare).astype(int), pts_per_part - floor)
        diff = num_points - target.sum()
        rounds = 0
        while diff != 0 and rounds < spare:
            grow = diff > 0
            ok = np.where(target < pts_per_part if grow else target > floor)[0]
            if len(ok) == 0:
                break
            w = coverage_per_part[ok]
            if w.sum() > 0:
                if not grow:
                    w = 1.0 / (w + 1e-8)
                pick = ok[np.random.choice(len(ok), p=w / w.sum())]
            else:
                pick = np.random.choice(ok)
            target[pick] += 1 if grow else -1
            diff += -1 if grow else 1
            rounds += 1
        if rounds >= spare:
            logger.warning("allocation fix-up hit its iteration limit, final diff %d", diff)
    return target


def allocate_fps_points(parts_data: Union[List[int], List[np.ndarray]], allocation_method: str, num_points: int,
                        min_points_per_part: int, voxel_size: float, voxel_ratio: float,
                        total_sample_points: Optional[List[int]] = None) -> np.ndarray:
    """Reference :87-148."""
    arrays = isinstance(parts_data[0], np.ndarray)
    sizes = np.array([len(p) for p in parts_data]) if arrays else np.array(parts_data)
    if allocation_method == "point_count":
        return _allocate_by_point_count_logic(sizes, num_points, min_points_per_part)
    if allocation_method == "spatial_coverage":
        if not arrays:
            raise ValueError("spatial_coverage allocation method requires point arrays, not point counts")
        coverage = np.array([calculate_voxel_coverage(p, voxel_size) for p in parts_data])
        return _allocate_by_spatial_coverage_logic(coverage, sizes, num_points, min_points_per_part)
    if allocation_method == "voxel_adaptive":
        if not arrays:
            raise ValueError("voxel_adaptive allocation method requires point arrays, not point counts")
        if total_sample_points is None:
            raise ValueError("voxel_adaptive allocation method requires total_sample_points parameter (per-part sample counts)")
        if not isinstance(total_sample_points, list):
            raise ValueError("voxel_adaptive allocation method requires total_sample_points to be a list of per-part counts")
        return np.minimum(np.array(total_sample_points), sizes)
    raise ValueError(f"Unknown allocation method: {allocation_method}")


def sample_farthest_points(points: torch.Tensor, lengths: Optional[torch.Tensor] = None, K: Union[int, List[int], torch.Tensor] = 50,
                           random_start_point: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """pytorch3d.ops.sample_farthest_points semantics (points (N,P,3) -> sampled (N,max K,3) zero padded, indices (N,max K) -1
    padded); with random_start_point the start of every cloud is one torch.randint draw from torch's global CPU generator."""
    if not points.is_cuda:
        raise RuntimeError("rap_b200 FPS needs CUDA tensors (no CPU fallback)")
    N, P, D = points.shape
    if D != 3:
        raise NotImplementedError("the FPS kernel is specialised for 3-D points")
    pts = points.detach().to(torch.float32).contiguous()
    lengths_h = torch.full((N,), P, dtype=torch.int64) if lengths is None else lengths.detach().cpu().to(torch.int64)
    if lengths_h.shape != (N,):
        raise ValueError("points and lengths must have same batch dimension.")
    if int(lengths_h.max()) > P:
        raise ValueError("A value in lengths was too large.")
    K_h = torch.full((N,), K, dtype=torch.int64) if isinstance(K, int) else torch.as_tensor(K).detach().cpu().to(torch.int64)
    if K_h.shape[0] != N:
        raise ValueError("K and points must have the same batch dimension")
    start = torch.zeros(N, dtype=torch.int64)
    if random_start_point:
        for n in range(N):
            start[n] = torch.randint(high=int(lengths_h[n]), size=(1,)).item()
    max_k = int(K_h.max())
    dev = pts.device
    idx = torch.empty(N, max_k, dtype=torch.int64, device=dev)
    out = torch.empty(N, max_k, 3, dtype=torch.float32, device=dev)
    l32, k32, s32 = (t.to(torch.int32).to(dev) for t in (lengths_h, K_h, start))
    check(lib.rap_fps_batched(ptr(pts), N, P, ptr(l32), ptr(k32), ptr(s32), max_k, ptr(idx), ptr(out), stream_ptr()), "rap_fps_batched")
    return out.to(points.dtype), idx


def apply_batched_fps(batch_augmented_tensor: torch.Tensor, batch_lengths_tensor: torch.Tensor, batch_k_tensor: torch.Tensor,
                      global_seed: int, device: torch.device) -> Tuple[List[torch.Tensor], torch.Tensor]:
    """Reference :263-309: seed torch, batched FPS with random start points, per-cloud sampled coordinates + the index matrix."""
    torch.manual_seed(global_seed)
    batch = batch_augmented_tensor.contiguous().to(device)
    _, indices = sample_farthest_points(batch, lengths=batch_lengths_tensor.to(device), K=batch_k_tensor.to(device), random_start_point=True)
    sampled = [batch[i][indices[i][:int(k)]] for i, k in enumerate(batch_k_tensor)]
    return sampled, indices
