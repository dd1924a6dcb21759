"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the RAP local feature extractor (MiniSpinNet).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.  Every function cites the reference file:line it
follows (paths relative to /root/reference/dataset_process/utils/spinnet).  Pinned against the reference class executed verbatim
(oracle/make_golden.py -> tests/golden/spinnet_*.npz).  `pytorch3d.ops.ball_query` (0.7.8) is absent offline and restated from its
published semantics -- that third-party boundary is "parity unpinned".
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

RAD_N, ELE_N, AZI_N, PATCH, VOXEL_SAMPLE, DELTA = 3, 7, 20, 512, 10, 0.8


def ball_query_first_k(queries: torch.Tensor, points: torch.Tensor, K: int, radius: float):
    """First K points (index order) with squared distance < radius^2 per query; idx padded -1 (pytorch3d.ops.ball_query)."""
    d2 = ((queries[:, None, :] - points[None, :, :]) ** 2).sum(-1)
    inside = d2 < radius * radius
    rank = torch.cumsum(inside.long(), dim=-1) - 1
    take = inside & (rank < K)
    idx = torch.full((queries.shape[0], K), -1, dtype=torch.long)
    i, j = take.nonzero(as_tuple=True)
    idx[i, rank[i, j]] = j
    return idx


def select_patches(points: torch.Tensor, keypoints: torch.Tensor, des_r: float, perm: np.ndarray) -> torch.Tensor:
    """patch_embedder.py:93-139: shuffle the cloud (injected permutation), first-512-in-order ball query of radius des_r,
    empty slots filled with the keypoint itself.  -> (K, 512, 3)"""
    pts = points[torch.as_tensor(perm, dtype=torch.long)]
    idx = ball_query_first_k(keypoints, pts, PATCH, des_r)
    patches = pts[idx.clamp(min=0)]
    return torch.where((idx < 0)[..., None], keypoints[:, None, :].expand_as(patches), patches)


def normalize_patches(patches: torch.Tensor, des_r: float) -> torch.Tensor:
    """axis_align with is_aligned_to_global_z=True (patch_embedder.py:141-166): centre = LAST slot of the patch, R = I;
    then normalize (:185-188): / des_r."""
    return (patches - patches[:, -1:, :]) / des_r


def voxel_centres() -> torch.Tensor:
    """get_voxel_coordinate(radius=1, rad_n=3, azi_n=20, ele_n=7) (utils/common.py:387-393, s2_grid :213-227,
    change_coordinates :355-370): 3 radial shells at 1/6,3/6,5/6 x 7 elevation x 20 azimuth -> (420,3), order (rad, ele, azi)."""
    beta = np.linspace(0, np.pi, ELE_N, endpoint=False) + np.pi / ELE_N / 2
    alpha = np.linspace(0, 2 * np.pi, AZI_N, endpoint=False) + np.pi / AZI_N
    B, A = np.meshgrid(beta, alpha, indexing="ij")
    B, A = B.flatten(), A.flatten()
    s2 = np.stack([np.sin(B) * np.cos(A), np.sin(B) * np.sin(A), np.cos(B)], axis=1)            # (140,3)
    scale = (np.arange(RAD_N) / RAD_N + 1 / (2 * RAD_N)).reshape(RAD_N, 1, 1)
    return torch.from_numpy((scale * s2[None]).reshape(-1, 3)).float()


def sphere_query(patches: torch.Tensor) -> torch.Tensor:
    """SPT -> sphere_query (patch_embedder.py:168-183, utils/common.py:396-440): per voxel centre the first 10 patch points
    (slot order) within 0.8/3; missing slots are zeros; slot 0 is zeroed when its point is patch slot 0 (mask1 quirk :422-424).
    -> (K, 420, 10, 3)"""
    centres = voxel_centres()
    out = torch.zeros(patches.shape[0], centres.shape[0], VOXEL_SAMPLE, 3)
    for k in range(patches.shape[0]):
        idx = ball_query_first_k(centres, patches[k], VOXEL_SAMPLE, DELTA / RAD_N)
        g = patches[k][idx.clamp(min=0)]
        g = torch.where((idx < 0)[..., None], torch.zeros_like(g), g)
        g[:, 0] = torch.where((idx[:, 0] == 0)[:, None], torch.zeros(3), g[:, 0])
        out[k] = g
    return out


def var_to_invar(pts: torch.Tensor) -> torch.Tensor:
    """utils/common.py:443-469: points of azimuth bin a rotated by -a * 18 deg about z (angles2rotation_matrix :82-93)."""
    K = pts.shape[0]
    p = pts.view(K, RAD_N, ELE_N, AZI_N, VOXEL_SAMPLE, 3)
    ang = -torch.arange(AZI_N, dtype=torch.float64) * (2 * math.pi / AZI_N)
    R = torch.zeros(AZI_N, 3, 3, dtype=torch.float64)
    R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = torch.cos(ang), -torch.sin(ang), torch.sin(ang), torch.cos(ang), 1.0
    R = R.float()
    return torch.einsum("krenjc,ndc->krenjd", p, R).reshape(K, -1, VOXEL_SAMPLE, 3)


def _bn(x, sd, prefix, affine):
    shape = (1, -1) + (1,) * (x.dim() - 2)
    y = (x - sd[prefix + "running_mean"].view(shape)) / torch.sqrt(sd[prefix + "running_var"].view(shape) + 1e-5)
    if affine:
        y = y * sd[prefix + "weight"].view(shape) + sd[prefix + "bias"].view(shape)
    return y


def _pad_azimuth_circular_elevation_zero(x):
    """pad_image / pad_image_3d with kernel 3 (utils/common.py:230-275): last dim circular, second-to-last zero."""
    x = torch.cat([x[..., -1:], x, x[..., :1]], dim=-1)
    zeros = torch.zeros_like(x[..., :1, :])
    return torch.cat([zeros, x, zeros], dim=-2)


def descriptor_cnn(sd: dict, inv_pts: torch.Tensor) -> torch.Tensor:
    """patch_embedder.py:74-84 + patchnet.py:49-84: 1x1 conv 3->16 + BN + ReLU, max over the 10 samples, view (16,3,7,20),
    Conv3d 16->64 (radial collapses 3->1) + BN + ReLU, 6 x [Conv2d 3x3 + BN(no affine) + ReLU], Conv2d 32->32,
    pool_layer attention weights, weighted mean, L2 normalisation.  -> (K, 32)"""
    x = inv_pts.permute(0, 3, 1, 2)                                              # (K,3,420,10)
    x = F.relu(_bn(F.conv2d(x, sd["pnt_layer.0.weight"], sd["pnt_layer.0.bias"]), sd, "pnt_layer.1.", True))
    x = x.max(dim=3).values.view(-1, 16, RAD_N, ELE_N, AZI_N)
    x = F.conv3d(_pad_azimuth_circular_elevation_zero(x), sd["conv_net.ops.0.weight"], sd["conv_net.ops.0.bias"]).squeeze(2)
    x = F.relu(_bn(x, sd, "conv_net.ops.1.", False))
    for conv, bn in ((3, 4), (6, 7), (9, 10), (12, 13), (15, 16), (18, 19)):
        x = F.conv2d(_pad_azimuth_circular_elevation_zero(x), sd[f"conv_net.ops.{conv}.weight"], sd[f"conv_net.ops.{conv}.bias"])
        x = F.relu(_bn(x, sd, f"conv_net.ops.{bn}.", False))
    x = F.conv2d(_pad_azimuth_circular_elevation_zero(x), sd["conv_net.ops.21.weight"], sd["conv_net.ops.21.bias"])
    w = F.relu(_bn(F.conv2d(x, sd["pool_layer.0.weight"], sd["pool_layer.0.bias"]), sd, "pool_layer.1.", True))
    w = F.relu(_bn(F.conv2d(w, sd["pool_layer.3.weight"], sd["pool_layer.3.bias"]), sd, "pool_layer.4.", True))
    f = (x * w).mean(dim=(2, 3))
    return F.normalize(f, p=2, dim=1)


def extract_features(sd: dict, points: torch.Tensor, keypoints: torch.Tensor, des_r: float, perm: np.ndarray, chunk: int = 256):
    """MiniSpinNet.forward (patch_embedder.py:49-91) -> {'desc' (K,32), 'patches' (K,512,3), 'pnt_max' per-voxel features}."""
    patches = normalize_patches(select_patches(points, keypoints, des_r, perm), des_r)
    desc = []
    for s in range(0, patches.shape[0], chunk):
        desc.append(descriptor_cnn(sd, var_to_invar(sphere_query(patches[s:s + chunk]))))
    return {"desc": torch.cat(desc), "patches": patches}
