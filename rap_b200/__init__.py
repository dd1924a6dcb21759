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


def install_extractor(module: str = "dataset_process.extract_sample_features") -> None:
    """Replace the reference's `FeatureExtractor` class (MiniSpinNet wrapper) by the CUDA extractor, leaving the rest of the
    module (SampleProcessor, FPS, outlier removal, CLI) untouched.  No-op-safe: raises ImportError if the reference module is absent."""
    ref = importlib.import_module(module)
    mine = importlib.import_module("rap_b200.dataset_process.extract_sample_features")
    ref.FeatureExtractor = mine.FeatureExtractor
