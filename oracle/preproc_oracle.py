"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the keypoint-selection stage in front of the feature extractor (SURVEY.md 8(f)
rank 1).  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this file.

Pinned against the reference run verbatim (oracle/make_golden.py preproc -> tests/golden/preproc_*.npz):
  * voxel_down_sample            <- dataset_process/utils/dataset_utils.py:279-322 (voxel_down_sample_torch, CPU)
  * voxel_coverage, sample counts, allocation  <- dataset_process/utils/point_sampling_utils.py:11-260
Third-party algorithms absent offline, restated from their published behaviour -- **parity unpinned** for these two:
  * farthest_point_sampling      <- pytorch3d 0.7.8 sample_farthest_points (CPU kernel semantics; reference call site
                                    point_sampling_utils.py:263-309: random_start_point=True, start index = torch.randint per cloud)
  * statistical_outlier_mask     <- open3d 0.18 PointCloud::RemoveStatisticalOutliers (reference call site
                                    extract_sample_features.py:360-413)
"""
from __future__ import annotations

import numpy as np
import torch

F32 = np.float32


# ------------------------------------------------------------------------------------------------ voxel down-sampling
def voxel_down_sample(points: np.ndarray, voxel_size: float) -> np.ndarray:
    """Index of the point closest (in 1000 quantised steps) to its voxel's centre, one per occupied voxel, ordered by the
    reference's flattened voxel id (dataset_utils.py:279-322).  All arithmetic in fp32 in the reference's operation order.
    Quirks kept: v_size = grid.max() (not +1), so voxels at the upper faces alias with their neighbours' ids."""
    p = np.ascontiguousarray(points, dtype=F32)
    vs = F32(voxel_size)
    n = p.shape[0]
    offset = np.floor(p.min(axis=0) / vs).astype(np.int64)
    grid = np.floor(p / vs)
    center = (grid + F32(0.5)) * vs
    d = p - center
    sq = d * d
    dist = np.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2])
    q = (dist / dist.max() * F32(999)).astype(np.int64)
    g = grid.astype(np.int64) - offset
    v = int(g.max())
    gid = g[:, 0] + g[:, 1] * v + g[:, 2] * v * v
    dec = 10 ** len(str(n - 1))
    key = np.arange(n, dtype=np.int64) + q * dec
    order = np.lexsort((key, gid))                      # by voxel id, then by (quantised distance, index)
    head = np.ones(n, bool)
    head[1:] = gid[order][1:] != gid[order][:-1]
    return (key[order][head] % dec).astype(np.int64)


def voxel_coverage(points: np.ndarray, voxel_size: float) -> int:
    """Occupied voxel count (point_sampling_utils.py:11-31): floor(points / voxel_size) in the array's own dtype."""
    if len(points) == 0:
        return 0
    return len(np.unique(np.floor(points / voxel_size).astype(int), axis=0))


# ------------------------------------------------------------------------------------------------ allocation (host logic)
def adaptive_sample_counts(parts, voxel_size, voxel_ratio, min_points_per_part, max_sample_points):
    """point_sampling_utils.py:33-85."""
    out = []
    for pts in parts:
        if len(pts) == 0:
            out.append(0)
            continue
        c = int(voxel_coverage(pts, voxel_size) * voxel_ratio)
        out.append(min(max_sample_points, min(len(pts), max(min_points_per_part, c))))
    return out


# ------------------------------------------------------------------------------------------------ farthest point sampling
def farthest_point_sampling(points: np.ndarray, lengths: np.ndarray, K: np.ndarray, start_idxs: np.ndarray) -> np.ndarray:
    """pytorch3d sample_farthest_points CPU kernel: (N, max(K)) int64 indices, -1 padded.  Each step: squared distances to the last
    pick (fp32, diff = last - p, sum over x,y,z in order, no fma), running minimum, then the FIRST not-yet-picked point with the
    largest running minimum (strict >, starting from -1)."""
    pts = np.ascontiguousarray(points, dtype=F32)
    N = pts.shape[0]
    out = np.full((N, int(np.max(K)) if N else 0), -1, np.int64)
    for n in range(N):
        L, k_n = int(lengths[n]), int(min(K[n], lengths[n]))
        if k_n <= 0:
            continue
        p = pts[n, :L]
        mind = np.full(L, np.finfo(F32).max, F32)
        picked = np.zeros(L, bool)
        last = int(start_idxs[n])
        out[n, 0] = last
        picked[last] = True
        for k in range(1, k_n):
            d = p[last] - p
            sq = d * d
            d2 = (sq[:, 0] + sq[:, 1]) + sq[:, 2]
            mind = np.minimum(mind, d2)
            cand = np.where(picked, F32(-1), mind)
            last = int(np.argmax(cand))               # first maximum
            out[n, k] = last
            picked[last] = True
    return out


def fps_start_indices(lengths, seed: int) -> np.ndarray:
    """The reference seeds torch's global CPU generator and draws one torch.randint(high=length) per cloud
    (point_sampling_utils.py:283, pytorch3d sample_farthest_points random_start_point=True)."""
    torch.manual_seed(seed)
    return np.array([int(torch.randint(high=int(L), size=(1,)).item()) for L in lengths], np.int64)


# ------------------------------------------------------------------------------------------------ statistical outlier removal
def statistical_outlier_mask(points: np.ndarray, nb_neighbors: int, std_ratio: float) -> np.ndarray:
    """open3d RemoveStatisticalOutliers: per point the mean of the distances (double) to its nb_neighbors nearest points (the point
    itself included, distance 0), threshold = mean + std_ratio * sample std (n-1) of those means; keep mean < threshold."""
    p = np.asarray(points, dtype=np.float64)
    n = len(p)
    avg = np.empty(n)
    k = min(nb_neighbors, n)
    for s in range(0, n, 1024):
        d2 = ((p[s:s + 1024, None, :] - p[None, :, :]) ** 2).sum(-1)
        part = np.sort(np.partition(d2, k - 1, axis=1)[:, :k], axis=1)
        avg[s:s + 1024] = np.sqrt(part).sum(1) / k
    mean = avg.sum() / n
    std = np.sqrt(((avg - mean) ** 2).sum() / (n - 1))
    return (avg > 0) & (avg < mean + std_ratio * std)
