"""Library-kernel comparators for bench.py (SURVEY.md 8(d) "(b)"): the kernels the reference's GPU path runs on a B200 --
FlashAttention-2 (`flash_attn_varlen_qkvpacked_func`, reference flow_model/layer.py:106-111,123-128), cuBLAS behind `nn.Linear` under
fp16 autocast (sample.py:23-25, config/trainer/infer.yaml:6) -- plus the other attention libraries present in the image (cuDNN SDPA,
flashinfer's Blackwell FMHA), timed on the EXACT shapes of the headline workload, in the same process, with CUDA events.

None of this is on the product path: these are the numbers the product's own kernels are compared with.  Nothing here imports `oracle/`;
`library_dit_forward` restates the DiT forward only to drive the library kernels in the order the reference does.
"""
from __future__ import annotations

import math
import os
import subprocess
import sys
import json

import torch
import torch.nn.functional as F


def _time(fn, iters: int = 20, warmup: int = 3) -> float:
    """Mean milliseconds per call (CUDA events on the current stream, after warm-up, synchronised on both sides)."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def attention_flops(cu: list[int], heads: int = 8, hd: int = 64) -> float:
    return sum(4.0 * heads * hd * (e - s) ** 2 for s, e in zip(cu[:-1], cu[1:]))


def attention_comparators(T: int, segmentations: dict[str, list[int]], heads: int = 8, hd: int = 64, iters: int = 20) -> dict:
    """TFLOP/s of each library attention on a (T,3,heads,hd) fp16 qkv tensor for every segmentation {name: cu_seqlens}."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(T, 3, heads, hd, device=dev, dtype=torch.float16, generator=g)
    out: dict = {}
    for name, cu in segmentations.items():
        fl = attention_flops(cu, heads, hd)
        res: dict = {"flops": fl}
        cu_t = torch.tensor(cu, dtype=torch.int32, device=dev)
        max_len = max(e - s for s, e in zip(cu[:-1], cu[1:]))
        # (i) FlashAttention-2, the reference's own call
        try:
            from flash_attn import flash_attn_varlen_qkvpacked_func
            ms = _time(lambda: flash_attn_varlen_qkvpacked_func(qkv, cu_t, max_len, 0.0, softmax_scale=1.0 / math.sqrt(hd)), iters)
            res["fa2_ms"], res["fa2_tflops"] = ms, fl / ms / 1e9
        except Exception as e:  # pragma: no cover
            res["fa2_error"] = repr(e)[:160]
        # (ii) cuDNN / torch-flash SDPA: equal-length segments become the batch dimension
        lens = {e - s for s, e in zip(cu[:-1], cu[1:])}
        if len(lens) == 1:
            n_seg, L = len(cu) - 1, lens.pop()
            q, k, v = (qkv[:, i].reshape(n_seg, L, heads, hd).transpose(1, 2) for i in range(3))
            from torch.nn.attention import SDPBackend, sdpa_kernel
            for key, backend in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("torch_flash", SDPBackend.FLASH_ATTENTION)):
                try:
                    with sdpa_kernel(backend):
                        ms = _time(lambda: F.scaled_dot_product_attention(q, k, v), iters)
                    res[f"{key}_ms"], res[f"{key}_tflops"] = ms, fl / ms / 1e9
                except Exception as e:  # pragma: no cover
                    res[f"{key}_error"] = repr(e)[:160]
        out[name] = res
    return out


def flashinfer_comparator(T: int, segmentations: dict[str, list[int]], heads: int = 8, hd: int = 64, timeout_s: float = 150.0) -> dict:
    """flashinfer's CUTLASS Blackwell FMHA (`flashinfer.prefill.fmha_varlen`).  The module is JIT-compiled on first use (nvcc on the box),
    so it runs in a subprocess under a timeout; a timeout / build failure is reported, not fatal."""
    payload = json.dumps({"T": T, "seg": segmentations, "heads": heads, "hd": hd})
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--flashinfer", payload], capture_output=True, text=True, timeout=timeout_s)
        for line in reversed(r.stdout.splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"error": (r.stderr or r.stdout)[-300:]}
    except subprocess.TimeoutExpired:
        return {"error": f"flashinfer JIT build did not finish within {timeout_s:.0f} s"}
    except Exception as e:  # pragma: no cover
        return {"error": repr(e)[:200]}


def _flashinfer_child(payload: str) -> None:
    cfg = json.loads(payload)
    T, heads, hd = cfg["T"], cfg["heads"], cfg["hd"]
    from flashinfer.prefill import fmha_varlen
    dev = torch.device("cuda")
    qkv = torch.randn(T, 3, heads, hd, device=dev, dtype=torch.float16)
    q, k, v = (qkv[:, i].contiguous() for i in range(3))
    out = {}
    for name, cu in cfg["seg"].items():
        cu_t = torch.tensor(cu, dtype=torch.int32, device=dev)
        max_len = max(e - s for s, e in zip(cu[:-1], cu[1:]))
        ms = _time(lambda: fmha_varlen(q, k, v, cu_t, cu_t, max_qo_len=max_len, causal=False, sm_scale=1.0 / math.sqrt(hd)), 20)
        fl = attention_flops(cu, heads, hd)
        out[name] = {"flashinfer_ms": ms, "flashinfer_tflops": fl / ms / 1e9}
    print(json.dumps(out), flush=True)


def gemm_comparators(M: int, shapes: dict[str, tuple[int, int]], iters: int = 30) -> dict:
    """cuBLAS(Lt) fp16 `F.linear` (no epilogue fused: bias only) on the DiT's GEMM shapes {name: (K, N)}: TFLOP/s of the bare GEMM the
    reference launches before its separate elementwise kernels."""
    dev = torch.device("cuda")
    out = {}
    for name, (K, N) in shapes.items():
        a = torch.randn(M, K, device=dev, dtype=torch.float16)
        w = torch.randn(N, K, device=dev, dtype=torch.float16) / math.sqrt(K)
        b = torch.randn(N, device=dev, dtype=torch.float16)
        ms = _time(lambda: F.linear(a, w, b), iters)
        out[name] = {"cublas_ms": ms, "cublas_tflops": 2.0 * M * N * K / ms / 1e9}
    return out


# --------------------------------------------------------------------------------------------------------------------------------
# Whole DiT forward on library kernels, in the order and precision the reference's GPU path uses them: fp16 autocast Linear (cuBLAS),
# FlashAttention-2 on fp16 q/k/v, fp32 LayerNorm / residual stream / qk-norm inputs as autocast leaves them.
# --------------------------------------------------------------------------------------------------------------------------------
def _posenc(x: torch.Tensor, n: int = 10) -> torch.Tensor:
    outs = [x]
    for k in range(n):
        f = float(2.0 ** k)
        outs += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(outs, dim=-1)


def _sincos(t: torch.Tensor, channels: int = 256) -> torch.Tensor:
    half = channels // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    a = t[:, None] * f[None]
    return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)


@torch.no_grad()
def library_dit_forward(sd: dict, x, timesteps, cond, feats, scales, anchor, cu_batch: torch.Tensor, cu_part: torch.Tensor, heads: int = 8):
    """(T,3) velocity.  `sd`: fp32 CUDA state_dict with the reference's keys; cu_*: int32 CUDA tensors."""
    from flash_attn import flash_attn_varlen_qkvpacked_func
    L = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("transformer_layers."))
    lens_b = (cu_batch[1:] - cu_batch[:-1]).long()
    max_b, max_p = int(lens_b.max()), int((cu_part[1:] - cu_part[:-1]).max())
    rep = lambda v: v.repeat_interleave(lens_b, dim=0)  # noqa: E731
    with torch.autocast("cuda", dtype=torch.float16):
        e = torch.cat([_posenc(cond), _posenc(x), _posenc(rep(scales)[:, None]), feats], dim=-1)
        h = F.linear(e, sd["encoding_manager.emb_proj.weight"], sd["encoding_manager.emb_proj.bias"]).float()
        emb = sd["anchor_part_emb.weight"]
        h = h + torch.where(anchor[:, None], emb[1][None], emb[0][None])
        T, d = h.shape
        hd = d // heads

        def adaln(prefix, hh):
            a = F.linear(_sincos(timesteps), sd[prefix + "timestep_embedder.linear_1.weight"], sd[prefix + "timestep_embedder.linear_1.bias"])
            a = F.linear(F.silu(a), sd[prefix + "timestep_embedder.linear_2.weight"], sd[prefix + "timestep_embedder.linear_2.bias"])
            a = F.linear(F.silu(a), sd[prefix + "linear.weight"], sd[prefix + "linear.bias"])
            sc, sh = a.chunk(2, dim=-1)
            return F.layer_norm(hh, (d,)) * (1 + rep(sc)) + rep(sh)

        def attn(p, blk, xx, cu, max_len):
            qkv = F.linear(xx, sd[p + f"{blk}_qkv_proj.weight"]).reshape(T, 3, heads, hd)
            q, k, v = qkv.unbind(dim=1)
            q = F.normalize(q.float(), dim=-1) * sd[p + f"{blk}_q_norm.gamma"] * (hd ** 0.5)
            k = F.normalize(k.float(), dim=-1) * sd[p + f"{blk}_k_norm.gamma"] * (hd ** 0.5)
            packed = torch.stack([q.half(), k.half(), v.half()], dim=1)
            o = flash_attn_varlen_qkvpacked_func(packed, cu, max_len, 0.0, softmax_scale=1.0 / math.sqrt(hd)).reshape(T, d)
            return F.linear(o, sd[p + f"{blk}_out_proj.weight"], sd[p + f"{blk}_out_proj.bias"])

        for i in range(L):
            p = f"transformer_layers.{i}."
            h = h + attn(p, "self", adaln(p + "self_prenorm.", h), cu_part, max_p).float()
            h = h + attn(p, "global", adaln(p + "global_prenorm.", h), cu_batch, max_b).float()
            xx = F.layer_norm(h, (d,), sd[p + "ff_norm.weight"], sd[p + "ff_norm.bias"])
            hg = F.linear(xx, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"])
            a, g = hg.chunk(2, dim=-1)
            h = h + F.linear(a * F.gelu(g), sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"]).float()
    # the reference runs final_mlp outside autocast's fp16 (point_cloud_dit.py:183-184: .float(), TF32 matmul)
    y = F.silu(F.linear(h, sd["final_mlp.0.weight"], sd["final_mlp.0.bias"]))
    y = F.silu(F.linear(y, sd["final_mlp.2.weight"], sd["final_mlp.2.bias"]))
    return F.linear(y, sd["final_mlp.4.weight"])


def library_forward_ms(sd: dict, batch, iters: int = 5) -> float:
    """Milliseconds per DiT forward of `batch` (device tensors) on the library-kernel stack, TF32 on like the reference (sample.py:23-25)."""
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    try:
        B = batch.cu_seqlens_batch.numel() - 1
        ts = torch.full((B,), 0.5, device=batch.cond.device)
        cu_b, cu_p = batch.cu_seqlens_batch.to(torch.int32), batch.cu_seqlens_part.to(torch.int32)
        return _time(lambda: library_dit_forward(sd, batch.x_1, ts, batch.cond, batch.features, batch.scales, batch.anchor_indices, cu_b, cu_p), iters, 2)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--flashinfer":
        _flashinfer_child(sys.argv[2])
