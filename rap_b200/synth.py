"""Deterministic synthetic inputs for tests and benchmarks (no dataset / checkpoint offline).

* `make_state_dict`  -- seeded random-init weights keyed exactly like the reference
  `PointCloudDiT.state_dict()` (keys verified against the reference class, see
  SURVEY.md 8(a) row W0 and reference `flow_model/point_cloud_dit.py:82-117`).
  Each tensor is drawn from its own generator seeded by (seed, key) so the result does
  not depend on module construction order and can be loaded into the reference class,
  the oracle and the CUDA path alike.
* `make_scene` / `make_batch` -- seeded multi-view scenes normalised the way the
  reference dataset does it (`data/dataset.py:756-793,804,867`): shift by the primary
  view's centroid, divide by 1.5*max|primary|, re-centre, every view individually
  centred to form the condition cloud, anchor view = largest (first on ties) keeps its
  GT placement.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class DiTConfig:
    embed_dim: int = 512
    num_layers: int = 12
    num_heads: int = 8
    local_feat_dim: int = 32
    multires: int = 10
    ada_channels: int = 256

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads

    @property
    def embed_in(self) -> int:  # 63 + 63 + 21 + feat
        m = self.multires
        return 2 * (3 + 6 * m) + (1 + 2 * m) + self.local_feat_dim


def state_dict_spec(cfg: DiTConfig) -> list[tuple[str, tuple[int, ...], str]]:
    """(key, shape, kind) in reference order; kind in {linear_w, head_w, linear_b, ones, ln_w, emb}."""
    d, h, hd = cfg.embed_dim, cfg.num_heads, cfg.head_dim
    spec = [("anchor_part_emb.weight", (2, d), "emb"),
            ("encoding_manager.emb_proj.weight", (d, cfg.embed_in), "linear_w"),
            ("encoding_manager.emb_proj.bias", (d,), "linear_b")]
    for i in range(cfg.num_layers):
        p = f"transformer_layers.{i}."
        for blk in ("self", "global"):
            spec += [(p + f"{blk}_prenorm.timestep_embedder.linear_1.weight", (d, cfg.ada_channels), "linear_w"),
                     (p + f"{blk}_prenorm.timestep_embedder.linear_1.bias", (d,), "linear_b"),
                     (p + f"{blk}_prenorm.timestep_embedder.linear_2.weight", (d, d), "linear_w"),
                     (p + f"{blk}_prenorm.timestep_embedder.linear_2.bias", (d,), "linear_b"),
                     (p + f"{blk}_prenorm.linear.weight", (2 * d, d), "linear_w"),
                     (p + f"{blk}_prenorm.linear.bias", (2 * d,), "linear_b"),
                     (p + f"{blk}_qkv_proj.weight", (3 * d, d), "linear_w"),
                     (p + f"{blk}_out_proj.weight", (d, d), "linear_w"),
                     (p + f"{blk}_out_proj.bias", (d,), "linear_b"),
                     (p + f"{blk}_q_norm.gamma", (h, hd), "ones"),
                     (p + f"{blk}_k_norm.gamma", (h, hd), "ones")]
        spec += [(p + "ff_norm.weight", (d,), "ln_w"),
                 (p + "ff_norm.bias", (d,), "linear_b"),
                 (p + "ff.net.0.proj.weight", (8 * d, d), "linear_w"),
                 (p + "ff.net.0.proj.bias", (8 * d,), "linear_b"),
                 (p + "ff.net.2.weight", (d, 4 * d), "linear_w"),
                 (p + "ff.net.2.bias", (d,), "linear_b")]
    spec += [("final_mlp.0.weight", (d, d), "linear_w"), ("final_mlp.0.bias", (d,), "linear_b"),
             ("final_mlp.2.weight", (d // 2, d), "linear_w"), ("final_mlp.2.bias", (d // 2,), "linear_b"),
             ("final_mlp.4.weight", (3, d // 2), "head_w")]
    return spec


def make_state_dict(cfg: DiTConfig = DiTConfig(), seed: int = 0) -> dict[str, torch.Tensor]:
    """Seeded fp32 weights with nn.Linear-like magnitudes (uniform +-1/sqrt(fan_in)).

    gamma / LayerNorm weights are 1 + 0.1*U(-1,1) and biases are non-zero so that every
    parameter of the reference participates in parity checks.
    """
    out: dict[str, torch.Tensor] = {}
    for key, shape, kind in state_dict_spec(cfg):
        g = torch.Generator(device="cpu")
        g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
        u = torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1
        if kind == "linear_w":
            t = u / (shape[-1] ** 0.5)
        elif kind == "head_w":      # O(1) velocities, like a trained v ~ x_1 - x_0
            t = u * 8.0 / (shape[-1] ** 0.5)
        elif kind == "linear_b":
            t = u * 0.05
        elif kind in ("ones", "ln_w"):
            t = 1.0 + 0.1 * u
        elif kind == "emb":
            t = u * 0.5
        else:  # pragma: no cover
            raise ValueError(kind)
        out[key] = t.contiguous()
    return out


# ----------------------------------------------------------------------------
# scenes
# ----------------------------------------------------------------------------
def _random_rotation(rng: np.random.Generator, roll_pitch_deg: float = 10.0) -> np.ndarray:
    yaw = rng.uniform(0, 2 * np.pi)
    roll, pitch = np.deg2rad(rng.uniform(-roll_pitch_deg, roll_pitch_deg, size=2))
    cz, sz, cy, sy, cx, sx = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    return rz @ ry @ rx


def _surface_cloud(rng: np.random.Generator, n: int) -> np.ndarray:
    """Union of random planar patches and boxes in a 10 m cube, 5 mm noise."""
    pts = []
    n_planes, n_boxes = int(rng.integers(6, 13)), int(rng.integers(2, 5))
    per = n // (n_planes + n_boxes) + 1
    for _ in range(n_planes):
        c = rng.uniform(-5, 5, 3)
        r = _random_rotation(rng, 90.0)
        uv = rng.uniform(-1, 1, (per, 2)) * rng.uniform(1.0, 4.0, 2)
        pts.append(c + np.c_[uv, np.zeros(per)] @ r.T)
    for _ in range(n_boxes):
        c = rng.uniform(-4, 4, 3)
        half = rng.uniform(0.3, 1.5, 3)
        p = rng.uniform(-1, 1, (per, 3)) * half
        face = rng.integers(0, 3, per)
        sign = rng.choice([-1.0, 1.0], per)
        p[np.arange(per), face] = sign * half[face]
        pts.append(c + p)
    pts = np.concatenate(pts)[:n]
    return pts + rng.normal(0, 0.005, pts.shape)


@dataclass
class Scene:
    cond: np.ndarray            # (V*N, 3) f32 per-view centred condition cloud (anchor view: GT placement)
    gt: np.ndarray              # (V*N, 3) f32 normalised GT assembly
    features: np.ndarray        # (V*N, F) f32 unit-norm
    scale: np.float32
    anchor: np.ndarray          # (V*N,) bool
    points_per_view: np.ndarray  # (V,) int64


def make_scene(scene_idx: int, views: int, pts_per_view: int | list[int], feat_dim: int = 32,
               overlap: float = 0.5, seed: int = 42) -> Scene:
    rng = np.random.default_rng(seed + scene_idx)
    counts = [pts_per_view] * views if isinstance(pts_per_view, int) else list(pts_per_view)
    assert len(counts) == views
    world = _surface_cloud(rng, max(40 * max(counts), 20000))
    parts = []
    for v in range(views):
        # half-space "frustum": keep the fraction of the world in front of a random plane so that
        # adjacent views overlap by roughly `overlap`.
        nrm = rng.normal(size=3)
        nrm /= np.linalg.norm(nrm)
        d = world @ nrm
        keep = d > np.quantile(d, 1.0 - (0.5 + 0.5 * overlap))
        idx = np.flatnonzero(keep)
        idx = rng.choice(idx, size=counts[v], replace=idx.size < counts[v])
        parts.append(world[idx])
    cnt = np.array(counts)
    offs = np.concatenate([[0], np.cumsum(cnt)])
    pts = np.concatenate(parts)
    primary = int(np.argmax(cnt))                                   # dataset.py:761 (first on ties)
    pp = pts[offs[primary]:offs[primary + 1]]
    primary_trans = pp.mean(0)
    scale = np.max(np.abs(pp - primary_trans)) * 1.5                # dataset.py:780
    gt = (pts - primary_trans) / scale                              # dataset.py:788,791
    gt_trans = gt.mean(0)
    gt = gt - gt_trans                                              # dataset.py:793
    cond = gt.copy()
    anchor = np.zeros(len(pts), bool)
    for v in range(views):
        st, ed = offs[v], offs[v + 1]
        order = rng.permutation(ed - st)                            # dataset.py:817
        gt[st:ed] = gt[st:ed][order]
        if v == primary:
            cond[st:ed] = gt[st:ed] + gt_trans                      # dataset.py:867
            anchor[st:ed] = True
        else:
            part = gt[st:ed]
            cond[st:ed] = part - part.mean(0)                       # dataset.py:804 (no rotation at test)
    feats = rng.standard_normal((len(pts), feat_dim))
    feats /= np.linalg.norm(feats, axis=1, keepdims=True)
    return Scene(cond.astype(np.float32), gt.astype(np.float32), feats.astype(np.float32),
                 np.float32(scale), anchor, cnt.astype(np.int64))


@dataclass
class Batch:
    """Packed batch as `variable_collate_fn` + `_prepare_data` produce it
    (reference `data/datamodule.py:169-198`, `modeling.py:203-231`)."""
    cond: torch.Tensor              # (T,3) f32
    features: torch.Tensor          # (T,F) f32
    scales: torch.Tensor            # (B,) f32
    anchor_indices: torch.Tensor    # (T,) bool
    points_per_part: torch.Tensor   # (B,P) i64 (zero padded)
    cu_seqlens_batch: torch.Tensor  # (B+1,) i32
    cu_seqlens_part: torch.Tensor   # (VP+1,) i32
    x_1: torch.Tensor               # (T,3) f32 injected noise
    gt: torch.Tensor                # (T,3) f32

    def to(self, device, non_blocking: bool = False) -> "Batch":
        return Batch(**{k: v.to(device, non_blocking=non_blocking) for k, v in self.__dict__.items()})


def make_batch(num_scenes: int, views: int, pts_per_view: int | list[int], max_parts: int | None = None,
               first_scene: int = 0, feat_dim: int = 32, overlap: float = 0.5) -> Batch:
    scenes = [make_scene(first_scene + i, views, pts_per_view, feat_dim, overlap) for i in range(num_scenes)]
    P = max_parts or views
    ppp = torch.zeros(num_scenes, P, dtype=torch.int64)
    for b, s in enumerate(scenes):
        ppp[b, :views] = torch.from_numpy(s.points_per_view)
    lens = ppp.sum(1)
    cu_b = torch.zeros(num_scenes + 1, dtype=torch.int32)
    cu_b[1:] = torch.cumsum(lens, 0)
    valid = ppp[ppp > 0]
    cu_p = torch.zeros(valid.numel() + 1, dtype=torch.int32)
    cu_p[1:] = torch.cumsum(valid, 0)
    T = int(lens.sum())
    x1 = torch.cat([torch.randn(int(lens[b]), 3, generator=torch.Generator("cpu").manual_seed(1234 + first_scene + b))
                    for b in range(num_scenes)])
    assert x1.shape[0] == T
    return Batch(
        cond=torch.from_numpy(np.concatenate([s.cond for s in scenes])),
        features=torch.from_numpy(np.concatenate([s.features for s in scenes])),
        scales=torch.tensor([float(s.scale) for s in scenes], dtype=torch.float32),
        anchor_indices=torch.from_numpy(np.concatenate([s.anchor for s in scenes])),
        points_per_part=ppp, cu_seqlens_batch=cu_b, cu_seqlens_part=cu_p, x_1=x1,
        gt=torch.from_numpy(np.concatenate([s.gt for s in scenes])))


# ----------------------------------------------------------------------------
# MiniSpinNet (local feature extractor) weights
# ----------------------------------------------------------------------------
def spinnet_state_dict_spec() -> list[tuple[str, tuple[int, ...], str]]:
    """(key, shape, kind) of the reference MiniSpinNet.state_dict() (dataset_process/utils/spinnet/patch_embedder.py:31-47,
    patchnet.py:68-84), verified against the reference class.  `Desc.`-prefixed in BUFFER-X style checkpoints."""
    def bn(prefix, c, affine):
        out = [(prefix + "weight", (c,), "ln_w"), (prefix + "bias", (c,), "linear_b")] if affine else []
        return out + [(prefix + "running_mean", (c,), "bn_mean"), (prefix + "running_var", (c,), "bn_var"), (prefix + "num_batches_tracked", (), "count")]
    spec = [("pnt_layer.0.weight", (16, 3, 1, 1), "conv_w"), ("pnt_layer.0.bias", (16,), "linear_b")] + bn("pnt_layer.1.", 16, True)
    spec += [("pool_layer.0.weight", (16, 32, 1, 1), "conv_w"), ("pool_layer.0.bias", (16,), "linear_b")] + bn("pool_layer.1.", 16, True)
    spec += [("pool_layer.3.weight", (1, 16, 1, 1), "conv_w"), ("pool_layer.3.bias", (1,), "pool_b")] + bn("pool_layer.4.", 1, True)
    spec += [("conv_net.ops.0.weight", (64, 16, 3, 3, 3), "conv_w"), ("conv_net.ops.0.bias", (64,), "linear_b")] + bn("conv_net.ops.1.", 64, False)
    chans = [(3, 64, 64), (6, 128, 64), (9, 128, 128), (12, 64, 128), (15, 64, 64), (18, 32, 64)]
    for i, co, ci in chans:
        spec += [(f"conv_net.ops.{i}.weight", (co, ci, 3, 3), "conv_w"), (f"conv_net.ops.{i}.bias", (co,), "linear_b")] + bn(f"conv_net.ops.{i + 1}.", co, False)
    spec += [("conv_net.ops.21.weight", (32, 32, 3, 3), "conv_w"), ("conv_net.ops.21.bias", (32,), "linear_b")]
    return spec


def make_spinnet_state_dict(seed: int = 0) -> dict[str, torch.Tensor]:
    """Seeded MiniSpinNet weights with non-trivial BatchNorm running statistics (eval mode uses them)."""
    out = {}
    for key, shape, kind in spinnet_state_dict_spec():
        g = torch.Generator(device="cpu")
        g.manual_seed((seed * 1000003 + zlib.crc32(("spinnet." + key).encode())) & 0x7FFFFFFF)
        if kind == "count":
            out[key] = torch.tensor(100, dtype=torch.long)
            continue
        u = torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        out[key] = {"conv_w": u * (1.5 / fan_in ** 0.5), "linear_b": u * 0.05, "pool_b": u * 0.05 + 0.5, "ln_w": 1.0 + 0.1 * u,
                    "bn_mean": u * 0.1, "bn_var": 0.6 + 0.3 * u}[kind].contiguous()
    return out


def make_raw_cloud(idx: int, n_points: int, n_keypoints: int, seed: int = 77):
    """A raw (un-normalised, metric) view cloud + keypoints drawn from it, for the extractor."""
    rng = np.random.default_rng(seed + idx)
    pts = _surface_cloud(rng, n_points).astype(np.float32)
    kp = pts[rng.choice(n_points, n_keypoints, replace=False)]
    return torch.from_numpy(pts), torch.from_numpy(kp)
