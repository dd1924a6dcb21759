"""`dataset_process.utils.dataset_utils` on the CUDA keypoint-selection kernels: only the voxel down-sampling the demo path uses
(reference dataset_utils.py:157-215, 279-322).  Results are the reference's CPU results index for index."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from ..._lib import check, lib, ptr, stream_ptr


def _cuda_f32(points) -> torch.Tensor:
    t = torch.as_tensor(points)
    if not torch.cuda.is_available():
        raise RuntimeError("rap_b200 keypoint selection needs a CUDA device (no CPU fallback)")
    return t.detach().to("cuda", torch.float32).contiguous()


def voxel_down_sample_torch(points: torch.Tensor, voxel_size: float) -> torch.Tensor:
    """(M,) int64 indices into `points`: per occupied voxel the point closest to the voxel centre (1000 quantised steps, ties to
    the lowest index), ordered by the reference's flattened voxel id (dataset_utils.py:279-322).  Deterministic (the reference
    flags its scatter_reduce as nondeterministic on CUDA, :317-319)."""
    p = _cuda_f32(points).reshape(-1, 3)
    n = p.shape[0]
    if n == 0:
        return torch.empty(0, dtype=torch.int64, device=p.device)
    out = torch.empty(n, dtype=torch.int64, device=p.device)
    count = C.c_int(0)
    check(lib.rap_voxel_downsample(ptr(p), n, float(voxel_size), ptr(out), C.byref(count), stream_ptr()), "rap_voxel_downsample")
    return out[:count.value]


def downsample_points(points: np.ndarray, normals: Optional[np.ndarray] = None, method: str = "voxel", voxel_size: float = 0.1,
                      num_points: Optional[int] = None, seed: Optional[int] = None,
                      use_torch: bool = True) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """dataset_utils.py:157-215 for method='voxel' (the one demo.py:277-283 uses); rows keep the input dtype."""
    if len(points) == 0:
        return points, normals
    if method != "voxel":
        raise NotImplementedError("rap_b200 implements the 'voxel' method of downsample_points (the demo path); use apply_batched_fps for FPS")
    idx = voxel_down_sample_torch(torch.from_numpy(np.ascontiguousarray(points)), voxel_size).cpu().numpy()
    return points[idx], (normals[idx] if normals is not None else None)
