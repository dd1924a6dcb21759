"""Per-kernel parity on the B200 (called through the C ABI): tcgen05 GEMM epilogues, varlen FMHA, Procrustes."""
import ctypes as C
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rap():
    from rap_b200 import _lib
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return _lib


def _sync_check(rap):
    torch.cuda.synchronize()
    assert rap.device_error() == 0


def _gemm(rap, epi, bn, A, W, bias=None, addend=None, gamma=None, out16_cols=None):
    M, K = A.shape
    N = W.shape[0]
    out32 = torch.zeros(M, N, device="cuda") if (epi & 0xff) == 0 else None
    out16 = torch.zeros(M, out16_cols or N, device="cuda", dtype=torch.float16) if ((epi & 0xff) != 0 or out16_cols) else None
    rap.check(rap.lib.rap_test_gemm(epi, bn, M, N, K, A.data_ptr(), W.data_ptr(), rap.ptr(bias), rap.ptr(addend), rap.ptr(out32),
                                    rap.ptr(out16), rap.ptr(gamma), None, rap.stream_ptr()), "rap_test_gemm")
    _sync_check(rap)
    return out32, out16


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (300, 512, 512, 128), (1024, 512, 2048, 128), (257, 512, 128, 256),
                                       (8192, 512, 512, 128)])
def test_gemm_f32_bias_residual(rap, M, N, K, bn):
    g = torch.Generator("cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(N, device="cuda", generator=g)
    add = torch.randn(M, N, device="cuda", generator=g)
    out32, _ = _gemm(rap, 0, bn, A, W, bias, add)
    ref = A.float() @ W.float().t() + bias + add
    err = (out32 - ref).abs().max().item()
    assert err < 2e-3, f"max err {err}"


@pytest.mark.parametrize("M,K,in_place,per_row", [(128, 64, False, False), (300, 512, True, True), (8192, 512, True, False), (1000, 2048, True, True),
                                                  (257, 128, False, True)])
def test_gemm_f32_layernorm_epilogue(rap, M, K, in_place, per_row):
    """EPI_F32_LN: the residual GEMM whose epilogue also writes LN(out) * (one_plus + mul[sel(row)]) + add[sel(row)] as the next GEMM's fp16
    operand (2-CTA clusters exchange the row moments).  in_place: addend aliases the output (the residual stream); per_row: rows select
    different modulation rows (several scenes in one tile), and a plain affine LayerNorm (one_plus = 0, stride 0) otherwise."""
    g = torch.Generator("cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = (torch.randn(512, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(512, device="cuda", generator=g)
    h = 3.0 * torch.randn(M, 512, device="cuda", generator=g) + 0.7          # non-zero mean: exercises the single-pass variance
    ref = A.float() @ W.float().t() + bias + h
    n_sel = 3 if per_row else 1
    tab = torch.randn(n_sel, 2, 512, device="cuda", generator=g)
    rows = (torch.arange(M, device="cuda") * n_sel // M).to(torch.int32) if per_row else torch.zeros(M, dtype=torch.int32, device="cuda")
    one_plus = 1.0 if per_row else 0.0
    out32 = h.clone() if in_place else torch.zeros(M, 512, device="cuda")
    addend = out32 if in_place else h
    ln_out = torch.zeros(M, 512, device="cuda", dtype=torch.float16)
    mul, add = tab[:, 0].contiguous(), tab[:, 1].contiguous()
    rap.check(rap.lib.rap_test_gemm_ln(M, K, A.data_ptr(), W.data_ptr(), bias.data_ptr(), addend.data_ptr(), out32.data_ptr(), mul.data_ptr(), add.data_ptr(),
                                       rows.data_ptr(), 512 if per_row else 0, one_plus, ln_out.data_ptr(), rap.stream_ptr()), "rap_test_gemm_ln")
    _sync_check(rap)
    assert (out32 - ref).abs().max().item() < 2e-3
    ln_ref = F.layer_norm(ref, (512,)) * (one_plus + mul[rows.long()]) + add[rows.long()]
    err = (ln_out.float() - ln_ref).abs().max().item()
    assert err < 2e-2, f"max err {err}"


@pytest.mark.parametrize("M,K,bias_on", [(8192, 512, True), (8192, 2048, True), (257, 128, True), (1000, 512, False), (128, 64, True), (40000, 512, True)])
def test_gemm_f32_bulk_reduction_epilogue(rap, M, K, bias_on):
    """EPI_F32_RED (5): the in-place residual GEMM h += A W^T + bias leaving through bulk tensor reductions (no residual read in the SM); ragged
    last row tile (rows past M dropped by the tensor map), several tiles per CTA (40000 rows: 626 tiles on 148 CTAs), twice in a row
    (the second launch accumulates on the first one's output)."""
    g = torch.Generator("cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g).half()
    W = (torch.randn(512, K, device="cuda", generator=g) / math.sqrt(K)).half()
    bias = torch.randn(512, device="cuda", generator=g) if bias_on else None
    h0 = torch.randn(M + 64, 512, device="cuda", generator=g)          # 64 guard rows behind the tensor: must stay untouched
    h = h0.clone()
    upd = A.float() @ W.float().t() + (bias if bias_on else 0)
    for rep in (1, 2):
        rap.check(rap.lib.rap_test_gemm(5, 256, M, 512, K, A.data_ptr(), W.data_ptr(), rap.ptr(bias), h.data_ptr(), h.data_ptr(), None, None, None,
                                        rap.stream_ptr()), "rap_test_gemm")
        _sync_check(rap)
        err = (h[:M] - (h0[:M] + rep * upd)).abs().max().item()
        assert err < 2e-3 * rep, f"pass {rep}: max err {err}"
        assert torch.equal(h[M:], h0[M:])
    # bit-identical to the register epilogue up to the order of the two additions: (acc + bias) + h here, (acc + h) + bias there
    ref32, _ = _gemm(rap, 0, 256, A, W, bias, h0[:M].contiguous())
    assert (ref32 + upd - h[:M]).abs().max().item() < 2e-3


@pytest.mark.parametrize("mode", [0x100, 0x300])
@pytest.mark.parametrize("M", [8192, 257, 1000])
def test_gemm_multicast_cluster_variant(rap, M, mode):
    """The opt-in pair GEMMs on 2-CTA clusters -- 0x100: weight tile fetched half by each CTA and multicast into both (two-arrival stage
    barriers); 0x300: `cta_group::2` UMMA, one 256x256 MMA per pair with each CTA holding half of the weight tile -- f32 residual and QKV +
    RMS-norm epilogues, even and odd numbers of 128-row tiles (257 -> a dummy partner tile)."""
    g = torch.Generator("cuda").manual_seed(M)
    A = torch.randn(M, 512, device="cuda", generator=g).half()
    W = (torch.randn(512, 512, device="cuda", generator=g) / 22.6).half()
    bias, add = torch.randn(512, device="cuda", generator=g), torch.randn(M, 512, device="cuda", generator=g)
    out32, _ = _gemm(rap, 0 | mode, 256, A, W, bias, add)
    assert (out32 - (A.float() @ W.float().t() + bias + add)).abs().max().item() < 2e-3
    Wq = (torch.randn(1536, 512, device="cuda", generator=g) / 22).half()
    gamma = 1 + 0.1 * torch.randn(2, 8, 64, device="cuda", generator=g)
    _, out16 = _gemm(rap, 2 | mode, 256, A, Wq, gamma=gamma)
    qkv = (A.float() @ Wq.float().t()).view(M, 3, 8, 64)
    ref = torch.stack([F.normalize(qkv[:, 0], dim=-1) * gamma[0] * 8, F.normalize(qkv[:, 1], dim=-1) * gamma[1] * 8, qkv[:, 2]], 1).reshape(M, 1536)
    assert (out16.float() - ref).abs().max().item() < 2e-2


def test_gemm_silu_f16(rap):
    g = torch.Generator("cuda").manual_seed(1)
    A = torch.randn(500, 512, device="cuda", generator=g).half()
    W = (torch.randn(256, 512, device="cuda", generator=g) / 22).half()
    bias = torch.randn(256, device="cuda", generator=g)
    _, out16 = _gemm(rap, 1, 128, A, W, bias)
    ref = F.silu(A.float() @ W.float().t() + bias)
    assert (out16.float() - ref).abs().max().item() < 1e-2


@pytest.mark.parametrize("scalar", [0, 1, 2])
def test_gemm_qkv_rmsnorm(rap, scalar):
    """scalar=0: the product's packed f32x2 epilogue; 1: the scalar form kept for A/B runs (epi | 0x400); 2: packed, tiles written by bulk tensor
    stores (epi | 0x800; 700 rows = a ragged last tile, clipped by the tensor map)."""
    g = torch.Generator("cuda").manual_seed(2)
    M = 700
    A = torch.randn(M, 512, device="cuda", generator=g).half()
    W = (torch.randn(1536, 512, device="cuda", generator=g) / 22).half()
    gamma = 1 + 0.1 * torch.randn(2, 8, 64, device="cuda", generator=g)
    _, out16 = _gemm(rap, 2 | {0: 0, 1: 0x400, 2: 0x800}[scalar], 256, A, W, gamma=gamma)
    qkv = (A.float() @ W.float().t()).view(M, 3, 8, 64)
    q = F.normalize(qkv[:, 0], dim=-1) * gamma[0] * 8
    k = F.normalize(qkv[:, 1], dim=-1) * gamma[1] * 8
    ref = torch.stack([q, k, qkv[:, 2]], 1).reshape(M, 1536)
    assert (out16.float() - ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("scalar", [0, 1, 2])
def test_gemm_geglu(rap, scalar):
    """GEGLU epilogue: h * gelu(g), exact-erf GELU.  scalar=0: the product's packed f32x2 form (erfc = 2^poly, one MUFU per output); 1: the
    scalar Abramowitz-Stegun form kept for A/B runs (epi | 0x400).  Gate values are scaled so that the tails (|g| up to ~9) are exercised."""
    g = torch.Generator("cuda").manual_seed(3)
    M = 333
    A = torch.randn(M, 512, device="cuda", generator=g).half()
    W = (torch.randn(4096, 512, device="cuda", generator=g) / 22).half()     # reference row order [h | g]
    W[2048:] *= 2.0                                                          # g ~ N(0, 2^2): reaches the clamp at |g| = 5.94
    bias = 0.1 * torch.randn(4096, device="cuda", generator=g)
    # pack rows the way rap_dit_load_weight does: per 256-row tile = 128 h rows then the matching 128 g rows
    idx = torch.cat([torch.cat([torch.arange(t * 128, t * 128 + 128), 2048 + torch.arange(t * 128, t * 128 + 128)]) for t in range(16)]).cuda()
    _, out16 = _gemm(rap, 3 | {0: 0, 1: 0x400, 2: 0x800}[scalar], 256, A, W[idx].contiguous(), bias[idx].contiguous(), out16_cols=2048)   # 2: bulk tensor stores
    hg = A.double() @ W.double().t() + bias
    ref = hg[:, :2048] * F.gelu(hg[:, 2048:])
    err = (out16.double() - ref).abs()
    # fp16 rounding of the output (2^-11 relative) + the fp32 accumulation; the GELU approximations themselves are below 1e-5 relative
    assert (err <= 6e-4 * ref.abs() + 2e-5).all(), f"max abs err {err.max().item():.3e}, worst ratio {(err / (6e-4 * ref.abs() + 2e-5)).max().item():.2f}"
    assert hg[:, 2048:].abs().max().item() > 6.5       # the clamp region is covered


# force=3: the kernel the product launches (one item per CTA), 2: with its last wave force-split; 0 / 1: the opt-in persistent stream-K kernel
# (wave-major schedule; 1 = a tiny per-CTA share, so small inputs are cut into pieces and recombined too); 4 / 5: the same with one
# contiguous range per CTA
@pytest.mark.parametrize("force", [3, 2, 7, 6, 8, 9, 0, 1, 4, 5])      # 6 / 7: as 2 / 3 with one UMMA-issuing warp per query tile; 8 / 9: the
# three-tile kernel (384 query rows per CTA on 64-key blocks), plain / last wave force-split
@pytest.mark.parametrize("lens", [[128], [200, 136, 72], [4096], [1, 127, 129, 300], [1024, 1024, 2048], [512, 512], [2000, 1500],
                                  [4096, 4096], [8192], [16384]])
def test_fmha_varlen(rap, lens, force):
    g = torch.Generator("cuda").manual_seed(sum(lens))
    T = sum(lens)
    qkv = torch.randn(T, 3, 8, 64, device="cuda", generator=g)
    qkv[:, :2] = F.normalize(qkv[:, :2], dim=-1) * 8        # q,k as the RMS norm leaves them
    qkv = qkv.half().contiguous()
    out = torch.zeros(T, 512, device="cuda", dtype=torch.float16)
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    rap.check(rap.lib.rap_test_fmha(qkv.data_ptr(), T, 8, rap.i32_array(cu), len(lens), force, out.data_ptr(), None, rap.stream_ptr()), "rap_test_fmha")
    _sync_check(rap)
    ref = torch.empty(T, 8, 64, device="cuda")
    for s, e in zip(cu[:-1], cu[1:]):
        q, k, v = (qkv[s:e, i].float().transpose(0, 1) for i in range(3))
        ref[s:e] = F.scaled_dot_product_attention(q[None], k[None], v[None])[0].transpose(0, 1)
    err = (out.float().view(T, 8, 64) - ref).abs().max().item()
    assert err < 1e-2, f"max err {err}"


@pytest.mark.parametrize("force", [2, 1, 9])
def test_fmha_lazy_reference_overflow_path(rap, force):
    """Adversarial scores for the lazy reference exponent: the first K/V block scores ~0, later blocks reach 2^40 times more, so
    the stale-reference pass must detect fp16 overflow of P and take the re-reference + O/l rescale path (several times)."""
    g = torch.Generator("cuda").manual_seed(11)
    T = 1024
    qkv = torch.randn(T, 3, 8, 64, device="cuda", generator=g)
    qkv[:, :2] = F.normalize(qkv[:, :2], dim=-1) * 16            # |q| = |k| = 16: scaled scores up to 46 in log2 units
    qkv[:128, 1] *= 1e-3                                         # block 0 keys ~ 0 -> reference exponent 0
    qkv[128:384, 1] *= 0.3                                       # intermediate growth
    qkv = qkv.half().contiguous()
    out = torch.zeros(T, 512, device="cuda", dtype=torch.float16)
    rap.check(rap.lib.rap_test_fmha(qkv.data_ptr(), T, 8, rap.i32_array([0, T]), 1, force, out.data_ptr(), None, rap.stream_ptr()), "rap_test_fmha")
    _sync_check(rap)
    q, k, v = (qkv[:, i].float().transpose(0, 1) for i in range(3))
    ref = F.scaled_dot_product_attention(q[None], k[None], v[None])[0].transpose(0, 1)
    assert torch.isfinite(out.float()).all()
    assert (out.float().view(T, 8, 64) - ref).abs().max().item() < 2e-2


def test_procrustes_kat(rap, golden_dir):
    from rap_b200.rectified_point_flow import procrustes
    g = np.load(os.path.join(golden_dir, "procrustes_kat.npz"))
    for i in range(5):
        src, tgt = torch.from_numpy(g[f"src{i}"]).cuda(), torch.from_numpy(g[f"tgt{i}"]).cuda()
        R, t = procrustes.solve_procrustes(src, tgt)
        _sync_check(rap)
        np.testing.assert_allclose(R.cpu().numpy(), g[f"R{i}"], atol=2e-5)
        np.testing.assert_allclose(t.cpu().numpy(), g[f"t{i}"], atol=2e-5)


def test_fit_and_rigidify_match_reference_golden(rap, golden_dir):
    from oracle.make_golden import CASES
    from rap_b200.rectified_point_flow import procrustes
    from rap_b200.synth import make_batch
    name = "ragged_b2_v3_L2_s3_rigid"
    _, B, V, N, _, _ = CASES[name]
    b = make_batch(B, V, N, max_parts=V + 2)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    pred = torch.from_numpy(g["end_point_trajectory"][-1]).cuda()
    cond, ppp, cu = b.cond.cuda(), b.points_per_part.cuda(), b.cu_seqlens_batch.cuda()
    R, t = procrustes.fit_transformations(cond, pred, ppp, cu)
    rig = procrustes.rigidify_prediction_with_procrustes(pred, cond, ppp, cu)
    _sync_check(rap)
    assert R.shape == (B, V + 2, 3, 3) and t.shape == (B, V + 2, 3)
    np.testing.assert_allclose(rig.cpu().numpy(), g["rigidified"], atol=2e-5)
    reps = ppp.reshape(-1)
    applied = torch.einsum("nij,nj->ni", R.reshape(-1, 3, 3).repeat_interleave(reps, 0), cond) + t.reshape(-1, 3).repeat_interleave(reps, 0)
    np.testing.assert_allclose(applied.cpu().numpy(), g["rigidified"], atol=2e-5)
    empty = (ppp == 0)
    assert R[empty].abs().max() == 0 and t[empty].abs().max() == 0
    with pytest.raises(ValueError):
        procrustes.fit_transformations(cond, pred, ppp, None)
    # (B,N,3) fixed-batching input (utils/point_clouds.py:50-55)
    Rb, tb = procrustes.fit_transformations(cond.view(B, -1, 3), pred.view(B, -1, 3), ppp)
    assert torch.equal(Rb, R) and torch.equal(tb, t)


def test_procrustes_planted_full_size(rap):
    """Size-independent property at BASELINE sizes: a planted rigid motion of 2x4096-point views is recovered."""
    from rap_b200.rectified_point_flow import procrustes
    g = torch.Generator("cuda").manual_seed(5)
    src = torch.randn(2 * 4096, 3, device="cuda", generator=g)
    q, _ = torch.linalg.qr(torch.randn(2, 3, 3, device="cuda", generator=g))
    q = q * torch.sign(torch.det(q))[:, None, None]
    tt = torch.randn(2, 3, device="cuda", generator=g)
    tgt = torch.cat([src[:4096] @ q[0].t() + tt[0], src[4096:] @ q[1].t() + tt[1]])
    ppp = torch.tensor([[4096, 4096]], device="cuda")
    R, t = procrustes.fit_transformations(src, tgt, ppp, torch.tensor([0, 8192], device="cuda", dtype=torch.int32))
    _sync_check(rap)
    assert (R[0] - q).abs().max() < 1e-5 and (t[0] - tt).abs().max() < 1e-5


def test_rigidity_rmse_matches_reference_formula(rap):
    """compute_rigidity_rmse (eval/metrics.py:511-622) restated in torch vs the fused kernel, both averaging modes, padded slots."""
    from rap_b200.rectified_point_flow import procrustes
    from rap_b200.rectified_point_flow.eval.metrics import compute_rigidity_rmse, fit_and_rigidity_sse
    g = torch.Generator("cuda").manual_seed(21)
    ppp = torch.tensor([[300, 0, 212], [0, 0, 0], [128, 64, 8]], device="cuda")
    T = int(ppp.sum())
    src = torch.randn(T, 3, device="cuda", generator=g)
    tgt = src + 0.05 * torch.randn(T, 3, device="cuda", generator=g)
    cu_b = torch.tensor([0, 512, 512, 712], device="cuda", dtype=torch.int32)
    R, t = procrustes.fit_transformations(src, tgt, ppp, cu_b)
    scales = torch.tensor([2.0, 3.0, 0.5], device="cuda")
    R2, t2, sse = fit_and_rigidity_sse(src, tgt, ppp)
    assert torch.equal(R2, R) and torch.equal(t2, t)
    for per_part in (False, True):
        got = compute_rigidity_rmse(src, tgt, R, t, ppp, cu_b, scales, average_per_part=per_part)
        want, off = [], 0
        for b in range(3):
            errs = []
            for p in range(3):
                n = int(ppp[b, p])
                if n == 0: continue
                e = ((src[off:off + n] @ R[b, p].t() + t[b, p] - tgt[off:off + n]) ** 2).sum(1)
                errs.append(e); off += n
            if not errs: want.append(float("inf"))
            elif per_part: want.append(float(torch.stack([x.mean().sqrt() for x in errs]).mean()) * float(scales[b]))
            else: want.append(float(torch.cat(errs).mean().sqrt()) * float(scales[b]))
        _sync_check(rap)
        assert got[1] == float("inf")
        assert abs(got[0].item() - want[0]) < 1e-5 and abs(got[2].item() - want[2]) < 1e-5


def test_trajectory_rigidity_selection(rap):
    """modeling.py:462-518: per-step fit + RMSE averaged over steps, argmin over generations; vs the per-step mirror calls."""
    from rap_b200.rectified_point_flow import procrustes
    from rap_b200.rectified_point_flow.eval.metrics import compute_rigidity_rmse, select_generation, trajectory_rigidity_rmse
    g = torch.Generator("cuda").manual_seed(5)
    ppp = torch.tensor([[100, 156, 0], [64, 0, 192]], device="cuda")
    cu_b = torch.tensor([0, 256, 512], device="cuda", dtype=torch.int32)
    cond = torch.randn(512, 3, device="cuda", generator=g)
    scales = torch.tensor([1.5, 0.7], device="cuda")
    scores = []
    for noise in (0.2, 0.02, 0.1):
        traj = cond[None] + noise * torch.randn(4, 512, 3, device="cuda", generator=g)
        got = trajectory_rigidity_rmse(cond, traj, ppp, scales)
        want = torch.stack([compute_rigidity_rmse(cond, traj[i], *procrustes.fit_transformations(cond, traj[i], ppp, cu_b), ppp, cu_b, scales)
                            for i in range(4)]).mean(0)
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7)
        scores.append(got)
    _sync_check(rap)
    assert select_generation(torch.stack(scores)).tolist() == [1, 1]


def test_rigidity_rmse_matches_reference_golden(rap, golden_dir):
    """compute_rigidity_rmse run verbatim from the reference (oracle/make_golden.py eval) vs the fused kernel, incl. an empty object."""
    import os
    from oracle.make_golden import eval_case
    from rap_b200.rectified_point_flow.eval import compute_rigidity_rmse
    g = np.load(os.path.join(golden_dir, "eval.npz"))
    d = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in eval_case().items()}
    args = (d["cond"], d["pred"], d["rotations_pred"], d["translations_pred"], d["points_per_part"], d["cu_seqlens_batch"])
    for per_part in (False, True):
        got = compute_rigidity_rmse(*args, d["scales"], average_per_part=per_part).cpu().numpy()
        _sync_check(rap)
        np.testing.assert_allclose(got, g[f"rigidity_rmse_{int(per_part)}"], rtol=2e-5)
    np.testing.assert_allclose(compute_rigidity_rmse(*args).cpu().numpy(), g["rigidity_rmse_noscale"], rtol=2e-5)


def test_apply_transform_full_resolution(rap):
    """demo.py:1330-1340: pose relative to the first scan applied to a full-resolution cloud; fp64 and fp32 clouds, numpy and tensor."""
    from rap_b200.rectified_point_flow.eval import apply_transform, relative_to_reference
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    T = np.eye(4); T[:3, :3] = q * np.sign(np.linalg.det(q)); T[:3, 3] = [10.5, -3.25, 0.125]
    T = relative_to_reference([np.eye(4), T])[1]
    pts = rng.standard_normal((2_000_003, 3)) * 50
    want = pts @ T[:3, :3].T + T[:3, 3]
    got64 = apply_transform(pts, T)
    assert got64.dtype == np.float64 and np.abs(got64 - want).max() < 1e-11
    got32 = apply_transform(torch.from_numpy(pts.astype(np.float32)).cuda(), T)
    _sync_check(rap)
    assert got32.is_cuda and got32.dtype == torch.float32 and np.abs(got32.cpu().numpy() - want).max() < 5e-5
    with pytest.raises(RuntimeError):
        bad = T.copy(); bad[3, 0] = 0.1
        apply_transform(pts[:10], bad)


def test_chamfer_and_overlap_match_reference_golden(rap, golden_dir):
    """compute_cd / compute_overlap_ratio on the nearest-neighbour kernel vs the reference run verbatim (oracle/make_golden.py eval)."""
    import os
    from oracle.make_golden import eval_case, eval_gt_case
    from rap_b200.rectified_point_flow.eval import compute_cd, compute_overlap_ratio
    g = np.load(os.path.join(golden_dir, "eval.npz"))
    d = eval_case()
    gt = eval_gt_case(d)
    pred, cu = d["pred"].cuda(), gt["cu2"].cuda()
    cd = compute_cd(gt["gt"].cuda(), pred, cu)
    _sync_check(rap)
    np.testing.assert_allclose(cd.cpu().numpy(), g["cd"], rtol=1e-5)
    ov = compute_overlap_ratio(pred, gt["ppp2"].cuda(), cu, taus=[0.05, 0.15, 0.4]).cpu().numpy()
    n = np.diff(gt["cu2"].numpy())
    assert ov.shape == (3, 2) and (np.abs(ov - g["overlap"]) <= 1.5 / n[None, :]).all()      # one borderline point may flip (cdist vs direct form)
    # brute-force check at a size the golden does not cover, ragged segments, single-part object -> 0
    gen = torch.Generator("cuda").manual_seed(2)
    ppp = torch.tensor([[700, 324, 0], [500, 0, 0], [33, 64, 100]], device="cuda")
    cu3 = torch.tensor([0, 1024, 1524, 1721], device="cuda")
    pts = torch.rand(1721, 3, device="cuda", generator=gen)
    got = compute_overlap_ratio(pts, ppp, cu3, taus=0.06)[0].cpu().numpy()
    lab = torch.repeat_interleave(torch.arange(3, device="cuda").repeat(3), ppp.reshape(-1))
    want = []
    for a, b in zip(cu3[:-1].tolist(), cu3[1:].tolist()):
        dm = torch.cdist(pts[a:b].double(), pts[a:b].double())
        dm[lab[a:b, None] == lab[None, a:b]] = float("inf")
        want.append(float((dm.min(1).values <= 0.06).double().mean()) if len(set(lab[a:b].tolist())) > 1 else 0.0)
    assert np.abs(got - np.array(want)).max() < 3e-3 and got[1] == 0.0
