"""Pose output in the reference's wire format: the `*_transform.txt` 4x4 files of
`Evaluator._save_transformation_files` (reference eval/evaluator.py:383-497) and demo.py's "relative to the first scan, applied to
the full-resolution cloud" step (demo.py:1293-1357).  Host code, except `apply_transform`, which streams the full-resolution cloud
through one HBM-bound kernel."""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from ..._lib import check, lib, ptr, stream_ptr


def relative_transform_matrix(R_pred, t_pred, R_gt, t_gt, scale: float, global_rotation=None, global_translation=None) -> np.ndarray:
    """4x4 float32 that maps the ground-truth-posed part onto its predicted pose, in metres:
    R = R_pred R_gt^T,  t = t_pred*s - (t_gt*s) R^T-row-form, right-multiplied by inv([R_global | t_global]) when the sample carries a
    global pre-alignment (evaluator.py:444-478).  Arithmetic and dtypes follow the reference (float64 pieces, float32 4x4)."""
    R_pred, R_gt = np.asarray(R_pred, dtype=float), np.asarray(R_gt, dtype=float)
    t_pred_m, t_gt_m = np.asarray(t_pred, dtype=float) * scale, np.asarray(t_gt, dtype=float) * scale
    rel_T = R_gt @ R_pred.T
    M = np.eye(4, dtype=np.float32)
    M[:3, :3] = rel_T.T
    M[:3, 3] = t_pred_m - t_gt_m @ rel_T
    if global_rotation is not None and global_translation is not None:
        G = np.eye(4, dtype=np.float32)
        G[:3, :3] = np.asarray(global_rotation, dtype=float)
        G[:3, 3] = np.asarray(global_translation, dtype=float)
        M = M @ np.linalg.inv(G)
    return M


def format_transform(matrix: np.ndarray) -> str:
    """Four lines of four '%12.8f' fields separated by one blank (evaluator.py:486-488)."""
    return "".join(" ".join(f"{v:12.8f}" for v in row) + "\n" for row in matrix)


def _suffix(generation_idx: Union[int, str]) -> str:
    if isinstance(generation_idx, str):
        return generation_idx if generation_idx.startswith("generation") else f"generation_{generation_idx}"
    return f"generation{generation_idx:02d}"


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def save_transformation_files(data: Dict[str, Any], idx: int, sample_dir, dataset_name: str, sample_idx: int,
                              generation_idx: Union[int, str] = 0, rotations_pred=None, translations_pred=None, global_rotation=None,
                              global_translation=None) -> List[Path]:
    """One `{dataset}_sample{idx:05d}_{generationNN}_part{pid:02d}_transform.txt` per non-empty part of batch element `idx`
    (same arguments as the reference method; returns the paths written)."""
    if rotations_pred is None or translations_pred is None:
        return []
    R_pred, t_pred = _np(rotations_pred[idx]), _np(translations_pred[idx])
    R_gt, t_gt = _np(data["rotations"][idx]), _np(data["translations"][idx])
    counts = _np(data["points_per_part"][idx])
    scale = float(data["scales"][idx])
    G_R = _np(global_rotation) if global_rotation is not None else None
    G_t = _np(global_translation) if global_translation is not None else None
    written = []
    for pid in np.nonzero(counts > 0)[0]:
        pid = int(pid)
        if pid >= len(R_pred) or pid >= len(R_gt):
            continue
        M = relative_transform_matrix(R_pred[pid], t_pred[pid], R_gt[pid], t_gt[pid], scale, G_R, G_t)
        path = Path(sample_dir) / f"{dataset_name}_sample{sample_idx:05d}_{_suffix(generation_idx)}_part{pid:02d}_transform.txt"
        path.write_text(format_transform(M))
        written.append(path)
    return written


def load_transform(path) -> np.ndarray:
    return np.loadtxt(path)


def relative_to_reference(transforms: Sequence[np.ndarray]) -> List[np.ndarray]:
    """demo.py:1330-1336: every pose expressed relative to the first scan, inv(T_0) @ T_i."""
    if not len(transforms):
        return []
    inv0 = np.linalg.inv(np.asarray(transforms[0]))
    return [inv0 @ np.asarray(T) for T in transforms]


def apply_transform(points: Union[np.ndarray, torch.Tensor], transform: np.ndarray) -> Union[np.ndarray, torch.Tensor]:
    """R p + t over a full-resolution cloud (open3d `PointCloud.transform` for an affine 4x4): numpy in -> numpy out (same dtype,
    through the GPU), CUDA tensor in -> CUDA tensor out."""
    is_np = isinstance(points, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(points)) if is_np else points
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    if not torch.cuda.is_available():
        raise RuntimeError("rap_b200 needs a CUDA device (no CPU fallback)")
    d = t.detach().to("cuda").contiguous().reshape(-1, 3)
    out = torch.empty_like(d)
    T = np.ascontiguousarray(np.asarray(transform, dtype=np.float64).reshape(16))
    check(lib.rap_apply_transform(ptr(d), d.shape[0], int(d.dtype == torch.float64), T.ctypes.data_as(C.POINTER(C.c_double)), ptr(out), stream_ptr()),
          "rap_apply_transform")
    return out.cpu().numpy() if is_np else out
