"""`dataset_process.extract_sample_features.FeatureExtractor` on the CUDA extractor (reference
dataset_process/extract_sample_features.py:59-220 wrapping MiniSpinNet, dataset_process/utils/spinnet/patch_embedder.py).

Same constructor arguments and `extract_features(points, keypoints)` contract (numpy in -> numpy out, tensor in -> tensor out,
`{'features': (K,32), 'keypoints'}`).  Differences by design: errors are raised instead of being swallowed into all-zero
features (reference :206-220) -- there is no CPU fallback -- and the model is the packed weight handle of the C library, not an
nn.Module.  The point shuffle uses the global numpy RNG exactly like the reference (`np.random.choice(N, N, replace=False)`,
patch_embedder.py:99), so seeding numpy the same way gives the same patches.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Union

import numpy as np
import torch

from .._lib import check, lib, ptr, stream_ptr

_DEFAULTS = {"num_points_per_patch": 512, "rad_n": 3, "azi_n": 20, "ele_n": 7, "delta": 0.8, "voxel_sample": 10}


class FeatureExtractor:
    def __init__(self, model_config: Dict | None = None, des_r: float = 3.0, is_aligned_to_global_z: bool = True,
                 checkpoint_path: str | None = None, device: str = "auto", state_dict: Dict[str, torch.Tensor] | None = None):
        cfg = dict(_DEFAULTS)
        cfg.update(model_config or {})
        if cfg != _DEFAULTS:
            raise NotImplementedError(f"the CUDA extractor is specialised for the released miniSpinNet geometry {_DEFAULTS}, got {cfg}")
        if not is_aligned_to_global_z:
            raise NotImplementedError("is_aligned_to_global_z=False (learned z-axis) is not used by the released pipeline")
        if device == "cpu" or not torch.cuda.is_available():
            raise RuntimeError("rap_b200 FeatureExtractor needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device in ("auto", "cuda") else torch.device(device)
        self.model_config, self.des_r, self.is_aligned_to_global_z = cfg, des_r, is_aligned_to_global_z
        h = C.c_void_p()
        check(lib.rap_spinnet_create(C.byref(h)), "rap_spinnet_create")
        self._h = h
        if state_dict is None:
            if not checkpoint_path or not os.path.exists(checkpoint_path):
                raise FileNotFoundError("FeatureExtractor needs MiniSpinNet weights: pass checkpoint_path (Desc.* keys) or state_dict")
            raw = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
            state_dict = {k[5:]: v for k, v in raw.items() if k.startswith("Desc.")}     # reference :121-129
        self.load_state_dict(state_dict)

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        expected = {}
        for i in range(lib.rap_spinnet_num_keys(self._h)):
            n = C.c_int64()
            key = lib.rap_spinnet_key(self._h, i, C.byref(n)).decode()      # (call first: `d[f(n)] = n.value` would read n before f runs)
            expected[key] = n.value
        missing = sorted(set(expected) - set(sd))
        unexpected = sorted(k for k in set(sd) - set(expected) if not k.endswith("num_batches_tracked"))
        if missing or unexpected:
            raise KeyError(f"MiniSpinNet state_dict mismatch: missing={missing[:4]} unexpected={unexpected[:4]}")
        keep = []
        with torch.cuda.device(self.device):
            for k, numel in expected.items():
                w = sd[k].detach().to(self.device, torch.float32).contiguous()
                if w.numel() != numel:
                    raise ValueError(f"{k}: {w.numel()} elements, expected {numel}")
                keep.append(w)
                check(lib.rap_spinnet_load_weight(self._h, k.encode(), ptr(w), numel, stream_ptr()), f"load {k}")
            torch.cuda.current_stream().synchronize()
            check(lib.rap_spinnet_finalize(self._h), "rap_spinnet_finalize")

    def extract_features(self, points: Union[np.ndarray, torch.Tensor], keypoints: Optional[Union[np.ndarray, torch.Tensor]] = None,
                         perm: np.ndarray | None = None) -> Dict[str, Union[np.ndarray, torch.Tensor]]:
        is_tensor = isinstance(points, torch.Tensor)
        if (points.numel() if is_tensor else len(points)) == 0:
            empty = torch.tensor([]) if is_tensor else np.array([])
            return {"features": empty, "keypoints": empty}
        if keypoints is None:
            keypoints = points.clone() if is_tensor else points.copy()
        to_dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a))).float().reshape(-1, 3).to(self.device).contiguous()  # noqa: E731
        pts, kps = to_dev(points), to_dev(keypoints)
        n = pts.shape[0]
        if perm is None:
            perm = np.random.choice(n, n, replace=False)          # same draw as the reference (global numpy RNG)
        perm_d = torch.from_numpy(np.asarray(perm).astype(np.int32)).to(self.device)
        desc = torch.empty(kps.shape[0], 32, device=self.device)
        with torch.cuda.device(self.device):
            check(lib.rap_spinnet_extract(self._h, ptr(pts), n, ptr(perm_d), ptr(kps), kps.shape[0], float(self.des_r), ptr(desc), stream_ptr()),
                  "rap_spinnet_extract")
        return {"features": desc if is_tensor else desc.cpu().numpy(), "keypoints": keypoints}

    def __del__(self):
        if getattr(self, "_h", None):
            lib.rap_spinnet_destroy(self._h)
            self._h = None


def remove_statistical_outlier(points: np.ndarray, nb_neighbors: int, std_ratio: float) -> np.ndarray:
    """Inlier indices of open3d's `PointCloud.remove_statistical_outlier(nb_neighbors, std_ratio)` (mean distance to the
    nb_neighbors nearest points, the point itself included, below mean + std_ratio * std over the cloud) on the CUDA KNN kernel."""
    pts = torch.from_numpy(np.ascontiguousarray(points)).to("cuda", torch.float32).contiguous()
    keep = torch.empty(pts.shape[0], dtype=torch.uint8, device=pts.device)
    n_keep = C.c_int(0)
    check(lib.rap_statistical_outliers(ptr(pts), pts.shape[0], int(nb_neighbors), float(std_ratio), ptr(keep), C.byref(n_keep), stream_ptr()),
          "rap_statistical_outliers")
    return torch.nonzero(keep).flatten().cpu().numpy()


class SampleProcessor:
    """`dataset_process.extract_sample_features.SampleProcessor` (reference :223-620) with every stage on the GPU: statistical
    outlier removal -> pre-FPS cap -> keypoint budget per part -> batched FPS -> descriptors with the ORIGINAL part as context.
    Same constructor arguments, same numpy / torch RNG call sequence, same return value (one dict per part)."""

    def __init__(self, feature_extractor: FeatureExtractor, num_points: int = 5000, skip_point_sampling: bool = False,
                 remove_outliers: bool = True, outlier_nb_neighbors: int = 20, outlier_std_ratio: float = 2.0, min_points_per_part: int = 100,
                 max_points_per_part: int = 10000, global_seed: int = 42, allocation_method: str = "point_count", voxel_size: float = 1.0,
                 voxel_ratio: float = 0.1):
        if allocation_method not in ("point_count", "spatial_coverage", "voxel_adaptive"):
            raise ValueError(f"allocation_method must be 'point_count', 'spatial_coverage', or 'voxel_adaptive', got: {allocation_method}")
        self.feature_extractor, self.num_points, self.skip_point_sampling = feature_extractor, num_points, skip_point_sampling
        self.remove_outliers, self.outlier_nb_neighbors, self.outlier_std_ratio = remove_outliers, outlier_nb_neighbors, outlier_std_ratio
        self.min_points_per_part, self.max_points_per_part, self.global_seed = min_points_per_part, max_points_per_part, global_seed
        self.allocation_method, self.voxel_size, self.voxel_ratio = allocation_method, voxel_size, voxel_ratio

    def _set_sample_seed(self, sample_idx: int = 0):
        seed = self.global_seed + sample_idx * 1000
        np.random.seed(seed)
        torch.manual_seed(seed)

    def select_keypoints(self, parts_points, parts_normals=None):
        """Stages 2-3 of process_sample (reference :359-575): returns (fps_parts, target_per_part, indices (n_parts, max K) tensor)."""
        from .utils.point_sampling_utils import allocate_fps_points, apply_batched_fps, calculate_adaptive_sample_count_per_part
        fps_parts, fps_normals = list(parts_points), list(parts_normals or [None] * len(parts_points))
        if self.remove_outliers:
            for i, part in enumerate(fps_parts):
                if len(part) < self.outlier_nb_neighbors:
                    continue
                inliers = remove_statistical_outlier(part, self.outlier_nb_neighbors, self.outlier_std_ratio)
                if len(inliers):
                    fps_parts[i] = part[inliers]
                    fps_normals[i] = fps_normals[i][inliers] if fps_normals[i] is not None else None
        if self.max_points_per_part is not None:
            cap = 20 * self.max_points_per_part
            for i, part in enumerate(fps_parts):
                if len(part) > cap:
                    keep = np.random.choice(len(part), cap, replace=False)
                    fps_parts[i] = part[keep].copy()
                    fps_normals[i] = fps_normals[i][keep].copy() if fps_normals[i] is not None else None
        if self.allocation_method == "voxel_adaptive":
            counts = calculate_adaptive_sample_count_per_part(fps_parts, self.voxel_size, self.voxel_ratio, self.min_points_per_part,
                                                              self.max_points_per_part)
            target = allocate_fps_points(fps_parts, self.allocation_method, self.num_points, self.min_points_per_part, self.voxel_size,
                                         self.voxel_ratio, total_sample_points=counts)
        elif self.allocation_method == "spatial_coverage":
            target = allocate_fps_points(fps_parts, self.allocation_method, self.num_points, self.min_points_per_part, self.voxel_size,
                                         self.voxel_ratio)
        else:
            target = allocate_fps_points(np.array([len(p) for p in fps_parts]), self.allocation_method, self.num_points,
                                         self.min_points_per_part, self.voxel_size, self.voxel_ratio)
        rows = [i for i, k in enumerate(target) if k != 0]
        if not rows:
            return fps_parts, fps_normals, target, rows, None, None
        lengths = [len(fps_parts[i]) for i in rows]
        ks = [min(int(target[i]), len(fps_parts[i])) for i in rows]
        batch = torch.zeros(len(rows), max(lengths), 3)
        for r, i in enumerate(rows):
            batch[r, :lengths[r]] = torch.from_numpy(fps_parts[i]).float()
        sampled, indices = apply_batched_fps(batch, torch.tensor(lengths, dtype=torch.int64), torch.tensor(ks, dtype=torch.int64),
                                             self.global_seed, self.feature_extractor.device)
        return fps_parts, fps_normals, target, rows, sampled, indices

    def process_sample(self, parts_points, parts_normals=None):
        if not parts_points:
            return []
        normals_in = parts_normals or [None] * len(parts_points)
        kept = [(p.copy(), n.copy() if n is not None else None) for p, n in zip(parts_points, normals_in) if len(p)]
        if not kept:
            return {"sampled_points": np.array([]), "sampled_normals": None, "features": np.array([])}
        originals, normals = [p for p, _ in kept], [n for _, n in kept]
        dev = self.feature_extractor.device
        if self.skip_point_sampling:
            out = []
            for pts, nrm in zip(originals, normals):
                t = torch.from_numpy(pts).float().to(dev)
                feats = self.feature_extractor.extract_features(points=t, keypoints=t)["features"]
                out.append({"sampled_points": pts, "sampled_normals": nrm, "features": feats.cpu().numpy()})
            return out
        fps_parts, fps_normals, _, rows, sampled, indices = self.select_keypoints(originals, normals)
        if sampled is None:
            return []
        out = []
        for r, kp in enumerate(sampled):
            if kp.numel() == 0:
                continue
            # NB the reference pairs sampled part r with originals[r] (extract_sample_features.py:593), i.e. positions, not `rows`
            feats = self.feature_extractor.extract_features(points=torch.from_numpy(originals[r]).float().to(dev), keypoints=kp)["features"]
            nrm = fps_normals[rows[r]]
            sel = indices[r][:kp.shape[0]].cpu().numpy()
            out.append({"sampled_points": kp.cpu().numpy(), "sampled_normals": nrm[sel] if nrm is not None else None,
                        "features": feats.cpu().numpy()})
        return out
