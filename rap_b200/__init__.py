"""rap_b200 -- B200-native (sm_100a) flow-matching registration sampler behind the reference's
`rectified_point_flow` module surface (PRBonn/RAP hot path; see DESIGN.md / INTEGRATION.md).

    import rap_b200
    rap_b200.install()          # aliases rectified_point_flow.{sampler,procrustes,flow_model} to this package
    # ... the reference's sample.py / demo.py / app.py then run unchanged on the CUDA kernels

Importing the compute modules requires librap_b200.so (built by `__graft_entry__.build()`); there is
no CPU or PyTorch fallback.  `rap_b200.synth` (synthetic inputs) imports without the library.
"""
from __future__ import annotations

import importlib
import sys

__version__ = "0.1.0"
_ALIASES = ("sampler", "procrustes", "flow_model")


def install(package: str = "rectified_point_flow") -> None:
    """Make `import rectified_point_flow.sampler` (etc.) resolve to the rap_b200 implementations.

    If the reference package is importable its other modules (modeling, eval, data, ...) keep working and
    only the three hot-path modules are replaced; call this BEFORE `import rectified_point_flow.modeling`
    so that its `from .sampler import get_sampler` binds to the fused sampler.
    """
    mine = importlib.import_module("rap_b200.rectified_point_flow")
    try:
        parent = importlib.import_module(package)
    except Exception:                       # reference not installed: expose the mirror package itself
        sys.modules[package] = mine
        parent = mine
    for name in _ALIASES:
        mod = importlib.import_module(f"rap_b200.rectified_point_flow.{name}")
        sys.modules[f"{package}.{name}"] = mod
        setattr(parent, name, mod)
    pc = importlib.import_module("rap_b200.rectified_point_flow.utils.point_clouds")
    sys.modules.setdefault(f"{package}.utils.point_clouds", pc)


def install_extractor(module: str = "dataset_process.extract_sample_features", keypoint_selection: bool = True) -> None:
    """Replace the reference's `FeatureExtractor` class (MiniSpinNet wrapper) by the CUDA extractor and -- with keypoint_selection --
    `SampleProcessor` plus the helpers it and demo.py call (`voxel_down_sample_torch`, `calculate_voxel_coverage`, `apply_batched_fps`, ...)
    by the GPU keypoint selection.  Raises ImportError if the reference module is absent."""
    ref = importlib.import_module(module)
    mine = importlib.import_module("rap_b200.dataset_process.extract_sample_features")
    ref.FeatureExtractor = mine.FeatureExtractor
    if not keypoint_selection:
        return
    ref.SampleProcessor = mine.SampleProcessor
    root = module.rsplit(".", 1)[0]                      # 'dataset_process'
    for sub, names in (("utils.dataset_utils", ("voxel_down_sample_torch", "downsample_points")),
                       ("utils.point_sampling_utils", ("calculate_voxel_coverage", "calculate_adaptive_sample_count_per_part",
                                                       "allocate_fps_points", "apply_batched_fps"))):
        src = importlib.import_module(f"rap_b200.dataset_process.{sub}")
        for cand in (f"{root}.{sub}", sub):              # the reference imports these both as package members and as top-level `utils.*`
            try:
                dst = importlib.import_module(cand)
            except Exception:
                continue
            for n in names:
                setattr(dst, n, getattr(src, n))
    for n in ("calculate_adaptive_sample_count_per_part", "allocate_fps_points", "apply_batched_fps"):
        if hasattr(ref, n):                              # names bound by `from utils.point_sampling_utils import ...` (reference :53)
            setattr(ref, n, getattr(importlib.import_module("rap_b200.dataset_process.utils.point_sampling_utils"), n))


def install_eval(package: str = "rectified_point_flow") -> None:
    """Replace the metric functions of `<package>.eval.metrics` that have CUDA mirrors (rigidity RMSE, chamfer, overlap ratio, pose
    errors); the Evaluator class itself stays the reference's."""
    ref = importlib.import_module(f"{package}.eval.metrics")
    mine = importlib.import_module("rap_b200.rectified_point_flow.eval.metrics")
    names = ("compute_rigidity_rmse", "compute_cd", "compute_overlap_ratio", "compute_transform_errors", "compute_transform_errors_direct")
    targets = [ref] + [sys.modules[m] for m in (f"{package}.eval.evaluator", f"{package}.modeling") if m in sys.modules]
    for mod in targets:                                  # evaluator.py / modeling.py bind the names with `from .metrics import ...`
        for n in names:
            if mod is ref or hasattr(mod, n):
                setattr(mod, n, getattr(mine, n))
