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
