"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the RAP flow-matching sampler path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product (`rap_b200/`) never does and fails loudly if the
CUDA library is missing.

Every function restates one piece of the reference's algorithm in plain torch on the CPU
(fp32 by default, fp64 on request to calibrate tolerances) and cites the reference
file:line it follows (paths relative to /root/reference/rectified_point_flow).

Parity status: the restatement is pinned against the reference's *own source executed
verbatim* in the build container (`oracle/make_golden.py` -> `tests/golden/*.npz`,
checked by `tests/test_oracle_golden.py`).  The reference ships no tests or golden
vectors of its own, and the `diffusers` pieces (GEGLU FeedForward, sinusoidal
Timesteps, TimestepEmbedding) are restated from the published diffusers 0.33.0 semantics
because the wheel is absent offline -- that third-party boundary is "parity unpinned".
"""
from __future__ import annotations

import math
from typing import Callable

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# embedding front-end
# ----------------------------------------------------------------------------
def posenc(x: torch.Tensor, num_freqs: int = 10) -> torch.Tensor:
    """NeRF encoding [x, sin(2^0 x), cos(2^0 x), ..., sin(2^9 x), cos(2^9 x)] in blocks of
    `input_dims` (flow_model/embedding.py:31-58; freq bands 2**linspace(0,9,10))."""
    outs = [x]
    for k in range(num_freqs):
        f = float(2.0 ** k)
        outs += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(outs, dim=-1)


def repeat_by_cu_seqlens(x: torch.Tensor, cu: torch.Tensor) -> torch.Tensor:
    """utils/point_clouds.py:161-183."""
    lens = (cu[1:] - cu[:-1]).long().to(x.device)
    return x.repeat_interleave(lens, dim=0)


def encode(sd: dict, x: torch.Tensor, cond: torch.Tensor, feats: torch.Tensor, scales_tok: torch.Tensor,
           anchor: torch.Tensor) -> torch.Tensor:
    """cat[posenc(cond) 63 | posenc(x_t) 63 | posenc(scale) 21 | local feat] -> Linear(179,512)
    (flow_model/embedding.py:131-179), then + anchor_part_emb[1] on anchor tokens and [0] elsewhere
    (flow_model/point_cloud_dit.py:119-139)."""
    e = torch.cat([posenc(cond), posenc(x), posenc(scales_tok[:, None]), feats], dim=-1)
    h = F.linear(e, sd["encoding_manager.emb_proj.weight"], sd["encoding_manager.emb_proj.bias"])
    emb = sd["anchor_part_emb.weight"]
    return h + torch.where(anchor[:, None], emb[1][None], emb[0][None])


# ----------------------------------------------------------------------------
# AdaLN / attention / FF
# ----------------------------------------------------------------------------
def timestep_sincos(t: torch.Tensor, channels: int = 256) -> torch.Tensor:
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0) as built at
    flow_model/norm.py:50-52: [cos(t f_k) | sin(t f_k)], f_k = exp(-ln(1e4) k / 128)."""
    half = channels // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half).to(t.dtype)
    a = t[:, None] * f[None]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


def adaln_scale_shift(sd: dict, prefix: str, t: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """flow_model/norm.py:71-73: Linear(SiLU(linear_2(SiLU(linear_1(sincos(t)))))).chunk(2)."""
    e = timestep_sincos(t)
    e = F.linear(e, sd[prefix + "timestep_embedder.linear_1.weight"], sd[prefix + "timestep_embedder.linear_1.bias"])
    e = F.linear(F.silu(e), sd[prefix + "timestep_embedder.linear_2.weight"], sd[prefix + "timestep_embedder.linear_2.bias"])
    e = F.linear(F.silu(e), sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])
    return e.chunk(2, dim=-1)


def adaln(sd: dict, prefix: str, x: torch.Tensor, t: torch.Tensor, cu_batch: torch.Tensor) -> torch.Tensor:
    """flow_model/norm.py:60-76: LN(x) (no affine, eps 1e-5) * (1 + scale) + shift, per scene."""
    scale, shift = adaln_scale_shift(sd, prefix, t)
    scale, shift = repeat_by_cu_seqlens(scale, cu_batch), repeat_by_cu_seqlens(shift, cu_batch)
    return F.layer_norm(x, (x.shape[-1],)) * (1 + scale) + shift


def varlen_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu: torch.Tensor) -> torch.Tensor:
    """Exact softmax attention inside each cu_seqlens segment, scale 1/sqrt(hd), non-causal
    (flash_attn_varlen_qkvpacked_func semantics as called at flow_model/layer.py:106-111,123-128).
    q,k,v: (T,H,hd)."""
    T, H, hd = q.shape
    out = torch.empty_like(q)
    bounds = cu.tolist()
    for s, e in zip(bounds[:-1], bounds[1:]):
        if e == s:
            continue
        qs, ks, vs = (t[s:e].transpose(0, 1)[None] for t in (q, k, v))               # (1,H,n,hd)
        out[s:e] = F.scaled_dot_product_attention(qs, ks, vs, scale=1.0 / math.sqrt(hd))[0].transpose(0, 1)
    return out


def attention_block(sd: dict, p: str, blk: str, x: torch.Tensor, cu: torch.Tensor, heads: int) -> torch.Tensor:
    """flow_model/layer.py:98-131 with MultiHeadRMSNorm (flow_model/norm.py:33):
    q,k <- normalize(.,dim=-1) * gamma[h] * sqrt(hd)."""
    T, d = x.shape
    hd = d // heads
    qkv = F.linear(x, sd[p + f"{blk}_qkv_proj.weight"]).reshape(T, 3, heads, hd)
    q, k, v = qkv.unbind(dim=1)
    q = F.normalize(q, dim=-1) * sd[p + f"{blk}_q_norm.gamma"] * (hd ** 0.5)
    k = F.normalize(k, dim=-1) * sd[p + f"{blk}_k_norm.gamma"] * (hd ** 0.5)
    o = varlen_attention(q, k, v, cu).reshape(T, d)
    return F.linear(o, sd[p + f"{blk}_out_proj.weight"], sd[p + f"{blk}_out_proj.bias"])


def feed_forward(sd: dict, p: str, x: torch.Tensor) -> torch.Tensor:
    """diffusers FeedForward(dim, activation_fn='geglu') used at flow_model/layer.py:89,164:
    h,g = proj(x).chunk(2); (h * gelu_erf(g)) @ W2^T + b2."""
    hg = F.linear(x, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"])
    h, g = hg.chunk(2, dim=-1)
    return F.linear(h * F.gelu(g), sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"])


def dit_layer(sd: dict, i: int, h: torch.Tensor, t: torch.Tensor, cu_batch: torch.Tensor, cu_part: torch.Tensor,
              heads: int) -> torch.Tensor:
    """flow_model/layer.py:134-166."""
    p = f"transformer_layers.{i}."
    h = h + attention_block(sd, p, "self", adaln(sd, p + "self_prenorm.", h, t, cu_batch), cu_part, heads)
    h = h + attention_block(sd, p, "global", adaln(sd, p + "global_prenorm.", h, t, cu_batch), cu_batch, heads)
    x = F.layer_norm(h, (h.shape[-1],), sd[p + "ff_norm.weight"], sd[p + "ff_norm.bias"])
    return h + feed_forward(sd, p, x)


def final_mlp(sd: dict, h: torch.Tensor) -> torch.Tensor:
    """flow_model/point_cloud_dit.py:111-117,183-184 (SiLU activations, last layer no bias)."""
    y = F.silu(F.linear(h, sd["final_mlp.0.weight"], sd["final_mlp.0.bias"]))
    y = F.silu(F.linear(y, sd["final_mlp.2.weight"], sd["final_mlp.2.bias"]))
    return F.linear(y, sd["final_mlp.4.weight"])


def num_layers_of(sd: dict) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("transformer_layers."))


def dit_forward(sd: dict, x: torch.Tensor, timesteps: torch.Tensor, cond: torch.Tensor, feats: torch.Tensor,
                scales: torch.Tensor, anchor: torch.Tensor, cu_batch: torch.Tensor, cu_part: torch.Tensor,
                heads: int = 8, return_transformer_features: bool = False):
    """PointCloudDiT.forward (flow_model/point_cloud_dit.py:141-191)."""
    dt = x.dtype
    sd = {k: v.to(dt) for k, v in sd.items()} if next(iter(sd.values())).dtype != dt else sd
    h = encode(sd, x, cond.to(dt), feats.to(dt), repeat_by_cu_seqlens(scales.to(dt), cu_batch), anchor)
    for i in range(num_layers_of(sd)):
        h = dit_layer(sd, i, h, timesteps.to(dt), cu_batch, cu_part, heads)
    v = final_mlp(sd, h)
    if return_transformer_features:
        return {"velocity": v, "transformer_features": h}
    return v


# ----------------------------------------------------------------------------
# Procrustes
# ----------------------------------------------------------------------------
def solve_procrustes(src: torch.Tensor, tgt: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """procrustes.py:6-37 (Kabsch via SVD, det fix by flipping the last row of Vt)."""
    sm, tm = src.mean(0, keepdim=True), tgt.mean(0, keepdim=True)
    H = (src - sm).t() @ (tgt - tm)
    U, _, Vt = torch.linalg.svd(H)
    R = Vt.t() @ U.t()
    if torch.det(R) < 0:
        Vt = Vt.clone()
        Vt[-1, :] *= -1
        R = Vt.t() @ U.t()
    t = tm - sm @ R.t()
    return R, t.squeeze(0)


def part_offsets(points_per_part: torch.Tensor) -> list[tuple[int, int, int, int]]:
    """(b, p, start, end) of every non-empty part in packing order (utils/point_clouds.py:27-48)."""
    out, off = [], 0
    for b, row in enumerate(points_per_part.tolist()):
        for p, n in enumerate(row):
            if n > 0:
                out.append((b, p, off, off + n))
                off += n
    return out


def fit_transformations(src: torch.Tensor, tgt: torch.Tensor, points_per_part: torch.Tensor):
    """procrustes.py:40-84: zeros for empty part slots."""
    B, P = points_per_part.shape
    src, tgt = src.reshape(-1, 3), tgt.reshape(-1, 3)
    R = torch.zeros(B, P, 3, 3, dtype=src.dtype)
    t = torch.zeros(B, P, 3, dtype=src.dtype)
    for b, p, s, e in part_offsets(points_per_part):
        R[b, p], t[b, p] = solve_procrustes(src[s:e], tgt[s:e])
    return R, t


def rigidify(pred: torch.Tensor, cond: torch.Tensor, points_per_part: torch.Tensor) -> torch.Tensor:
    """procrustes.py:86-118: out[part] = cond_part @ R^T + t with (R,t) = procrustes(cond_part -> pred_part)."""
    out = torch.zeros_like(pred)
    for _, _, s, e in part_offsets(points_per_part):
        R, t = solve_procrustes(cond[s:e], pred[s:e])
        out[s:e] = cond[s:e] @ R.t() + t
    return out


# ----------------------------------------------------------------------------
# sampler
# ----------------------------------------------------------------------------
def flow_sampler(flow_model_fn: Callable[[torch.Tensor, float], torch.Tensor], x_1: torch.Tensor, cond: torch.Tensor,
                 points_per_part: torch.Tensor, num_steps: int, rigidity_forcing: bool):
    """sampler.py:11-92 with the Euler step: t = 1 - step*dt; v = f(x_t,t); x0_hat = x_t - v t;
    x_t <- x_t - dt v; optionally x_t <- rigid(x0_hat) (1 - t + dt) + x_1 (t - dt).
    Returns (end_point_trajectory, trajectory), each (S,T,3)."""
    dt = 1.0 / num_steps
    x_t = x_1.clone()
    ep, tr = [], []
    for step in range(num_steps):
        t = 1 - step * dt
        v = flow_model_fn(x_t, t)
        x0_hat = x_t - v * t
        x_t = x_t - dt * v
        if rigidity_forcing:
            x_t = rigidify(x0_hat, cond, points_per_part) * (1 - t + dt) + x_1 * (t - dt)
        ep.append(x0_hat.clone())
        tr.append(x_t.clone())
    return torch.stack(ep), torch.stack(tr)


def register(sd: dict, cond: torch.Tensor, feats: torch.Tensor, scales: torch.Tensor, anchor: torch.Tensor,
             points_per_part: torch.Tensor, cu_batch: torch.Tensor, cu_part: torch.Tensor, x_1: torch.Tensor,
             num_steps: int, rigidity_forcing: bool, heads: int = 8, dtype: torch.dtype = torch.float32):
    """The path `sample_rectified_flow` + `fit_transformations` takes in modeling.py:633-741,365-389:
    sampler -> prediction = last *un-rigidified* x0_hat -> per-part Procrustes cond -> prediction."""
    sd = {k: v.to(dtype) for k, v in sd.items()}
    cond_, x1_ = cond.to(dtype), x_1.to(dtype)
    B = cu_batch.numel() - 1

    def fn(x, t):
        return dit_forward(sd, x, torch.full((B,), t, dtype=dtype), cond_, feats, scales, anchor, cu_batch, cu_part, heads)

    ep, tr = flow_sampler(fn, x1_, cond_, points_per_part, num_steps, rigidity_forcing)
    R, t = fit_transformations(cond_, ep[-1], points_per_part)
    return {"end_point_trajectory": ep, "trajectory": tr, "rotations": R, "translations": t}
