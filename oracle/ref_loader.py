"""TEST INFRASTRUCTURE ONLY -- loads the reference's own hot-path modules verbatim.

This file is part of `oracle/`: only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.  It never
runs on the GPU box (there is no /root/reference there); it exists so that
`oracle/make_golden.py` can execute the *unmodified* reference sources in the
build container and pin the restatement in `oracle/rap_oracle.py` against them.

How: the reference package cannot be imported as-is because
`rectified_point_flow/utils/__init__.py:2` pulls `hydra` and
`flow_model/layer.py:6` / `norm.py:10` pull `diffusers` and `flash_attn`.  We
pre-seed `sys.modules` with empty package objects whose `__path__` points at
the reference directories (so the package `__init__` files never execute) and
register two third-party shims:

* `diffusers.models.attention.FeedForward`, `diffusers.models.embeddings.
  {Timesteps, TimestepEmbedding}` -- restated from the published diffusers
  0.33.0 semantics (pinned only in reference `scripts/install.sh:9`;
  **unpinned by any reference test** -- see DESIGN.md "parity status").
* `flash_attn.flash_attn_varlen_qkvpacked_func` -- SDPA per `cu_seqlens`
  segment, scale 1/sqrt(head_dim), non-causal (reference call sites
  `flow_model/layer.py:106-111,123-128`).

No reference source is copied; the files are executed where they lie.
"""
from __future__ import annotations

import importlib
import math
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("RAP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rectified_point_flow"))


# ----------------------------------------------------------------------------
# third-party shims (diffusers 0.33.0 semantics, flash_attn 2.x semantics)
# ----------------------------------------------------------------------------
class _Timesteps(nn.Module):
    """diffusers.models.embeddings.Timesteps (sinusoidal, no parameters)."""

    def __init__(self, num_channels: int, flip_sin_to_cos: bool, downscale_freq_shift: float, scale: int = 1):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift
        self.scale = scale

    def forward(self, timesteps: torch.Tensor) -> torch.Tensor:
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.downscale_freq_shift)
        emb = torch.exp(exponent)
        emb = timesteps[:, None].float() * emb[None, :]
        emb = self.scale * emb
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class _TimestepEmbedding(nn.Module):
    """diffusers.models.embeddings.TimestepEmbedding: linear_2(SiLU(linear_1(x)))."""

    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample: torch.Tensor) -> torch.Tensor:
        return self.linear_2(self.act(self.linear_1(sample)))


class _GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h, g = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(g)


class _FeedForward(nn.Module):
    """diffusers.models.attention.FeedForward(dim, activation_fn='geglu'), mult=4."""

    def __init__(self, dim: int, dropout: float = 0.0, activation_fn: str = "geglu"):
        super().__init__()
        if activation_fn != "geglu":
            raise NotImplementedError(activation_fn)
        inner = dim * 4
        self.net = nn.ModuleList([_GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim)])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for m in self.net:
            x = m(x)
        return x


def _flash_attn_varlen_qkvpacked_func(qkv, cu_seqlens, max_seqlen, dropout_p=0.0, softmax_scale=None,
                                      causal=False, window_size=(-1, -1), softcap=0.0, **_):
    """qkv (total, 3, H, D); exact softmax attention inside each cu_seqlens segment."""
    assert softcap == 0.0 and not causal
    total, _, H, D = qkv.shape
    scale = softmax_scale if softmax_scale is not None else 1.0 / math.sqrt(D)
    out = torch.empty(total, H, D, dtype=qkv.dtype, device=qkv.device)
    cu = cu_seqlens.tolist()
    for s, e in zip(cu[:-1], cu[1:]):
        if e == s:
            continue
        q, k, v = (qkv[s:e, i].transpose(0, 1) for i in range(3))  # (H, n, D)
        out[s:e] = F.scaled_dot_product_attention(q[None], k[None], v[None], scale=scale)[0].transpose(0, 1)
    return out


def _install_shims() -> None:
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "diffusers" not in sys.modules:
        mod("diffusers").__path__ = []
        mod("diffusers.models").__path__ = []
        mod("diffusers.models.attention", FeedForward=_FeedForward)
        mod("diffusers.models.embeddings", Timesteps=_Timesteps, TimestepEmbedding=_TimestepEmbedding)
    # Always shadow flash_attn on CPU: the real wheel is CUDA-only at run time.
    fa = types.ModuleType("flash_attn")
    fa.flash_attn_varlen_qkvpacked_func = _flash_attn_varlen_qkvpacked_func
    sys.modules["flash_attn"] = fa


_PKG = "rectified_point_flow"


def load_reference():
    """Returns a namespace with the reference's own `sampler`, `procrustes`,
    `point_clouds` and `flow_model` modules (executed verbatim)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_shims()
    root = os.path.join(REFERENCE_ROOT, _PKG)
    for name, path in ((_PKG, root), (_PKG + ".utils", os.path.join(root, "utils"))):
        if name not in sys.modules or not getattr(sys.modules[name], "_rap_bypass", False):
            m = types.ModuleType(name)
            m.__path__ = [path]
            m._rap_bypass = True
            sys.modules[name] = m
    ns = types.SimpleNamespace()
    ns.point_clouds = importlib.import_module(_PKG + ".utils.point_clouds")
    ns.procrustes = importlib.import_module(_PKG + ".procrustes")
    ns.sampler = importlib.import_module(_PKG + ".sampler")
    ns.flow_model = importlib.import_module(_PKG + ".flow_model")
    return ns


def load_reference_eval():
    """The reference's own `eval.metrics` (for `compute_rigidity_rmse`) and `eval.evaluator.Evaluator` (for the `*_transform.txt` writer,
    `_save_transformation_files`) executed verbatim.  lightning / open3d / pytorch3d / the renderer are only named at import time on
    these two code paths, so empty stand-ins are enough."""
    load_reference()

    def stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            m.__dict__.update(attrs)
            sys.modules[name] = m
        else:
            for k, v in attrs.items():
                if not hasattr(sys.modules[name], k):
                    setattr(sys.modules[name], k, v)
    stub("lightning", LightningModule=object)
    stub("open3d")
    stub("pytorch3d")
    stub("pytorch3d.loss")
    stub("pytorch3d.loss.chamfer", chamfer_distance=None)
    stub("pytorch3d.ops", iterative_closest_point=None)
    stub(_PKG + ".utils.render", part_ids_to_colors=None)
    root = os.path.join(REFERENCE_ROOT, _PKG, "eval")
    if _PKG + ".eval" not in sys.modules:
        m = types.ModuleType(_PKG + ".eval")
        m.__path__ = [root]
        sys.modules[_PKG + ".eval"] = m
    ns = types.SimpleNamespace()
    ns.metrics = importlib.import_module(_PKG + ".eval.metrics")
    ns.evaluator = importlib.import_module(_PKG + ".eval.evaluator")
    return ns


def load_reference_modeling(install_rap_b200: bool = True):
    """The reference's own `modeling.py` (class `RectifiedPointFlow`: `_prepare_data`, `sample_rectified_flow`, `test_step`) executed
    verbatim, with -- when `install_rap_b200` -- `rectified_point_flow.{sampler,procrustes,flow_model}` aliased to the rap_b200 mirrors
    BEFORE modeling.py runs its `from .sampler import get_sampler` / `from .procrustes import fit_transformations`, i.e. exactly what
    `rap_b200.install()` does for sample.py / demo.py.  `lightning` is absent offline: `LightningModule` becomes a plain `nn.Module`
    with the `device` property and no-op loggers (the only Lightning features these code paths touch)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    unload_reference()
    _install_shims()
    root = os.path.join(REFERENCE_ROOT, _PKG)
    for name, path in ((_PKG, root), (_PKG + ".utils", os.path.join(root, "utils")), (_PKG + ".eval", os.path.join(root, "eval"))):
        m = types.ModuleType(name)
        m.__path__ = [path]
        m._rap_bypass = True
        sys.modules[name] = m
    importlib.import_module(_PKG + ".utils.point_clouds")          # the reference's own glue (install() only fills it in if absent)
    if install_rap_b200:
        import rap_b200
        rap_b200.install(_PKG)
    else:
        for sub in ("procrustes", "sampler", "flow_model"):
            importlib.import_module(f"{_PKG}.{sub}")

    class LightningModule(nn.Module):
        @property
        def device(self):
            p = next(self.parameters(), None)
            return p.device if p is not None else torch.device("cpu")

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    L = types.ModuleType("lightning")
    L.__path__ = []
    L.LightningModule = LightningModule
    sys.modules["lightning"] = L
    for name, attrs in (("open3d", {}), ("pytorch3d", {}), ("pytorch3d.loss", {}), ("pytorch3d.loss.chamfer", {"chamfer_distance": None}),
                        ("pytorch3d.ops", {"iterative_closest_point": None}), (_PKG + ".utils.render", {"part_ids_to_colors": None})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            if not hasattr(sys.modules[name], k):
                setattr(sys.modules[name], k, v)
    ns = types.SimpleNamespace()
    ns.modeling = importlib.import_module(_PKG + ".modeling")
    ns.sampler = sys.modules[_PKG + ".sampler"]
    ns.procrustes = sys.modules[_PKG + ".procrustes"]
    ns.flow_model = sys.modules[_PKG + ".flow_model"]
    return ns


def unload_reference() -> None:
    """Remove the bypass packages so `rap_b200.install()` can be tested after."""
    for k in [k for k in sys.modules if k == _PKG or k.startswith(_PKG + ".")]:
        del sys.modules[k]
    if getattr(sys.modules.get("lightning"), "__path__", None) == []:      # our stand-in, not a real install
        del sys.modules["lightning"]


# ----------------------------------------------------------------------------
# local feature extractor (MiniSpinNet): reference dataset_process/utils/spinnet/*
# ----------------------------------------------------------------------------
def _ball_query_shim(p1, p2, lengths1=None, lengths2=None, K=500, radius=0.2, return_nn=True):
    """pytorch3d.ops.ball_query (0.7.8, pinned only in reference scripts/install.sh:16 -- absent offline, restated):
    for every query of p1 the FIRST K points of p2 (in index order) with squared distance < radius^2;
    idx padded with -1, dists (squared) and nn padded with 0."""
    B, P1, _ = p1.shape
    P2 = p2.shape[1]
    d2 = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)                   # (B,P1,P2)
    inside = d2 < radius * radius
    rank = torch.cumsum(inside.long(), dim=-1) - 1                                # position among the in-radius points
    take = inside & (rank < K)
    idx = torch.full((B, P1, K), -1, dtype=torch.long, device=p1.device)
    dists = torch.zeros(B, P1, K, dtype=p1.dtype, device=p1.device)
    b, i, j = take.nonzero(as_tuple=True)
    idx[b, i, rank[b, i, j]] = j
    dists[b, i, rank[b, i, j]] = d2[b, i, j]
    nn_pts = None
    if return_nn:
        nn_pts = torch.zeros(B, P1, K, p2.shape[2], dtype=p2.dtype, device=p2.device)
        nn_pts[b, i, rank[b, i, j]] = p2[b, j]
    return dists, idx, nn_pts


def load_reference_spinnet():
    """The reference's own `MiniSpinNet` class executed verbatim (CPU): package-bypass for `dataset_process`
    (its `utils/__init__.py` imports open3d/h5py tooling) + a `pytorch3d.ops.ball_query` shim.  The hard-coded `.cuda()`
    at patch_embedder.py:176 is neutralised by the caller (see `reference_spinnet_forward`)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    root = os.path.join(REFERENCE_ROOT, "dataset_process")
    for name, path in (("dataset_process", root), ("dataset_process.utils", os.path.join(root, "utils"))):
        if name not in sys.modules or not getattr(sys.modules[name], "_rap_bypass", False):
            m = types.ModuleType(name)
            m.__path__ = [path]
            m._rap_bypass = True
            sys.modules[name] = m
    if "pytorch3d" not in sys.modules:
        p3 = types.ModuleType("pytorch3d"); p3.__path__ = []
        ops = types.ModuleType("pytorch3d.ops"); ops.ball_query = _ball_query_shim
        sys.modules["pytorch3d"], sys.modules["pytorch3d.ops"] = p3, ops
    return importlib.import_module("dataset_process.utils.spinnet.patch_embedder").MiniSpinNet


def reference_spinnet_forward(model, pts, kpts, des_r, numpy_seed):
    """Runs the reference MiniSpinNet.forward on CPU with numpy seeded (the point shuffle at patch_embedder.py:99 uses the global
    numpy RNG) and `.cuda()` turned into a no-op for the duration of the call."""
    import numpy as np
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        np.random.seed(numpy_seed)
        with torch.no_grad():
            return model(pts=pts, kpts=kpts, des_r=des_r, is_aligned_to_global_z=True)
    finally:
        torch.Tensor.cuda = orig
