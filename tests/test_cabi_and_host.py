"""CPU-side checks: the C-ABI library loads and exports every symbol include/rap_b200.h declares, the host-side
Kabsch solve matches the oracle, layout helpers, module surface / drop-in aliasing, and loud failure without a GPU."""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "rap_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rap_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rap_b200 import _lib
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/rap_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert _lib.lib.rap_version() >= 100


def test_kabsch_host_matches_oracle_incl_reflection_and_rank_deficiency():
    from oracle import rap_oracle as orc
    from rap_b200 import _lib
    g = torch.Generator().manual_seed(3)
    for n, reflect in [(50, False), (50, True), (3, False), (1000, False)]:
        src = torch.randn(n, 3, generator=g, dtype=torch.float64)
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
        if (torch.det(q) < 0) != reflect:
            q[:, 0] *= -1
        tgt = src @ q.t() + 0.01 * torch.randn(n, 3, generator=g, dtype=torch.float64)
        H = ((src - src.mean(0)).t() @ (tgt - tgt.mean(0))).contiguous()
        Hc, Rc = (C.c_double * 9)(*H.flatten().tolist()), (C.c_double * 9)()
        assert _lib.lib.rap_kabsch_host(Hc, Rc) == 0
        R = torch.tensor(list(Rc), dtype=torch.float64).view(3, 3)
        Ro, _ = orc.solve_procrustes(src, tgt)
        assert abs(float(torch.det(R)) - 1) < 1e-9
        assert (R - Ro).abs().max() < 1e-8


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rap_b200 import _lib
    from rap_b200.rectified_point_flow import procrustes
    from rap_b200.rectified_point_flow.flow_model import PointCloudDiT
    h = C.c_void_p()
    rc = _lib.lib.rap_dit_create(C.byref(h), 512, 2, 8, 32)
    assert rc != 0 and b"no CUDA device" in _lib.lib.rap_last_error()
    with pytest.raises(RuntimeError, match="CUDA"):
        procrustes.solve_procrustes(torch.zeros(4, 3), torch.zeros(4, 3))
    m = PointCloudDiT(in_dim=0, out_dim=3, embed_dim=512, num_layers=1, num_heads=8, local_feat_dim=32)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(4, 3), torch.zeros(1), torch.zeros(4, 3), torch.zeros(4, 32), None, torch.ones(1),
          torch.zeros(4, dtype=torch.bool), torch.tensor([0, 4]), torch.tensor([0, 4]))


def test_state_dict_surface_matches_reference_keys():
    from rap_b200.rectified_point_flow.flow_model import PointCloudDiT
    from rap_b200.synth import DiTConfig, make_state_dict, state_dict_spec
    for L in (1, 3):
        m = PointCloudDiT(in_dim=0, out_dim=3, embed_dim=512, num_layers=L, num_heads=8, local_feat_dim=32)
        spec = state_dict_spec(DiTConfig(num_layers=L))
        sd = m.state_dict()
        assert list(sd) == [k for k, _, _ in spec]
        assert all(tuple(sd[k].shape) == s for k, s, _ in spec)
        missing, unexpected = m.load_state_dict(make_state_dict(DiTConfig(num_layers=L)), strict=True)
        assert not missing and not unexpected
    with pytest.raises(NotImplementedError):
        PointCloudDiT(in_dim=0, out_dim=3, embed_dim=256, num_layers=1, num_heads=8)


def test_sampler_surface_and_closure_introspection():
    from rap_b200.rectified_point_flow import sampler
    from rap_b200.rectified_point_flow.flow_model import PointCloudDiT
    with pytest.raises(ValueError, match="Unknown sampler"):
        sampler.get_sampler("rk2")
    f = sampler.get_sampler("euler")
    assert f.func is sampler.flow_sampler and f.args == (sampler.euler_step,)

    class Harness:
        flow_model = PointCloudDiT(in_dim=0, out_dim=3, embed_dim=512, num_layers=1, num_heads=8, local_feat_dim=32)

    self = Harness()
    cond = local_features = scales = anchor_indices = cu_seqlens_batch = cu_seqlens_part = torch.zeros(1)
    latent_features = None

    def closure(x, t):
        return self.flow_model, cond, local_features, latent_features, scales, anchor_indices, cu_seqlens_batch, cu_seqlens_part

    model, ctx, cells = sampler._fused_context(closure)
    assert model is self.flow_model and ctx["cond"] is cond and "self" in cells
    assert sampler._fused_context(lambda x, t: x) is None          # opaque callables take the per-step route


def test_install_aliases_reference_module_names():
    import rap_b200
    saved = {k: v for k, v in sys.modules.items() if k == "rectified_point_flow" or k.startswith("rectified_point_flow.")}
    for k in saved:
        del sys.modules[k]
    try:
        rap_b200.install()
        import rectified_point_flow.sampler as s
        from rectified_point_flow.flow_model import PointCloudDiT
        from rectified_point_flow.procrustes import fit_transformations
        assert s.__name__ == "rap_b200.rectified_point_flow.sampler"
        assert PointCloudDiT.__module__.startswith("rap_b200.") and fit_transformations.__module__.startswith("rap_b200.")
    finally:
        for k in [k for k in sys.modules if k == "rectified_point_flow" or k.startswith("rectified_point_flow.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_points_per_part_from_cu_roundtrip():
    from rap_b200.engine import points_per_part_from_cu
    ppp = np.array([[200, 136, 72, 0], [5, 0, 0, 0], [64, 64, 64, 64]])
    cu_b = np.concatenate([[0], np.cumsum(ppp.sum(1))])
    cu_p = np.concatenate([[0], np.cumsum(ppp[ppp > 0])])
    out = points_per_part_from_cu(cu_b, cu_p)
    assert out.shape == (3, 4) and (out == ppp).all()


def test_split_parts_and_repeat_helpers():
    from rap_b200.rectified_point_flow.utils.point_clouds import repeat_by_cu_seqlens, split_parts
    x = torch.arange(30.).view(10, 3)
    parts = split_parts(x, torch.tensor([[2, 0, 4], [4, 0, 0]]), torch.tensor([0, 6, 10]))
    assert [len(p) for p in parts] == [2, 1] and parts[0][1].shape == (4, 3) and torch.equal(parts[1][0], x[6:])
    with pytest.raises(ValueError):
        split_parts(x, torch.tensor([[10]]))
    r = repeat_by_cu_seqlens(torch.tensor([[1.], [2.]]), torch.tensor([0, 2, 5]))
    assert r.flatten().tolist() == [1, 1, 2, 2, 2]


def _fmha_schedule(cu, heads=8, force=1):
    import ctypes as C
    from rap_b200 import _lib
    cu_a = _lib.i32_array(cu)
    n = _lib.lib.rap_fmha_schedule(cu_a, len(cu) - 1, heads, force, None, 0, None, None)
    buf, ns, mp = (C.c_int32 * (8 * n))(), C.c_int(), C.c_int()
    assert _lib.lib.rap_fmha_schedule(cu_a, len(cu) - 1, heads, force, buf, n, C.byref(ns), C.byref(mp)) == n
    return np.array(buf).reshape(n, 8), ns.value, mp.value


@pytest.mark.parametrize("cu", [[0, 4096, 8192], [0, 8192], [0, 512, 1024], [0, 16384], [0, 3000, 5000, 9000], [0, 100], [0, 130, 700, 701]])
def test_fmha_schedule_covers_every_key_block_once(cu):
    """Host logic of the attention launch: whole units plus key-split pieces must tile (query block, head) x keys exactly, pieces of a
    unit share a slot and are numbered 0..n-1, and big pieces are issued before small ones."""
    items, n_slots, max_pieces = _fmha_schedule(cu)
    seen = {}
    for q0, s0, s1, head, lo, hi, slot, piece in items.tolist():
        assert s0 <= q0 < s1 and s0 <= lo < hi <= s1 and (lo - s0) % 128 == 0 and (hi == s1 or (hi - s0) % 128 == 0)
        n_p, idx = piece >> 8, piece & 0xff
        assert (slot < 0) == (n_p == 1) and idx < n_p <= max(max_pieces, 1)
        seen.setdefault((q0, head), []).append((lo, hi, slot, idx, n_p))
    q_blocks = sum(-(-(b - a) // 256) for a, b in zip(cu[:-1], cu[1:]))
    assert len(seen) == 8 * q_blocks
    slots = set()
    for (q0, head), pieces in seen.items():
        pieces.sort()
        seg = next((a, b) for a, b in zip(cu[:-1], cu[1:]) if a <= q0 < b)
        assert pieces[0][0] == seg[0] and pieces[-1][1] == seg[1] and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
        assert len({p[2] for p in pieces}) == 1 and sorted(p[3] for p in pieces) == list(range(pieces[0][4]))
        if pieces[0][2] >= 0:
            slots.add(pieces[0][2])
    assert slots == set(range(n_slots))
    # cfg2 (the headline layout): 148 whole units, then 108 big and 108 small pieces -> every SM ends together
    if cu in ([0, 4096, 8192], [0, 8192]):
        assert (items[:148, 6] < 0).all() and n_slots == 108 and max_pieces == 2
        blocks = (items[:, 5] - items[:, 4]) // 128
        assert (blocks[148:256] == blocks[148]).all() and (blocks[256:] == blocks[256]).all() and blocks[148] == 3 * blocks[256]
        # the cost model (piece overhead ~4.5 key blocks) keeps the 32-block per-view launch whole and cuts the 64-block per-scene one
        plain, n_plain, _ = _fmha_schedule(cu, force=0)
        assert (n_plain == 0 and len(plain) == 256) if cu == [0, 4096, 8192] else (n_plain == 108 and len(plain) == 364)


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected before any CUDA call (errno-style code + rap_last_error text), so this runs on the CPU box."""
    import ctypes as C
    from rap_b200 import _lib
    L = _lib.lib
    n = C.c_int(0)
    assert L.rap_voxel_downsample(None, 0, 0.5, None, C.byref(n), None) == -22 and b"rap_voxel_downsample" in L.rap_last_error()
    assert L.rap_voxel_coverage(None, -1, 0, 1.0, C.byref(n), None) == -22
    assert L.rap_voxel_coverage(None, 0, 0, 1.0, C.byref(n), None) == 0 and n.value == 0          # empty part: 0 voxels, no device work
    assert L.rap_fps_batched(None, 1, 10, None, None, None, 5, None, None, None) == -22
    assert L.rap_fps_capacity() >= 100_000                                                       # the reference's pre-FPS cap (20 x 5000)
    assert L.rap_statistical_outliers(None, 5, 20, 2.0, None, C.byref(n), None) == -22
    assert L.rap_segment_nn(None, None, None, 1, None, None, None, None) == -22
    T = (C.c_double * 16)(*([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0.5, 0, 0, 1]))                 # not affine
    assert L.rap_apply_transform(1, 10, 0, T, 1, None) == -22 and b"affine" in L.rap_last_error()
    assert L.rap_procrustes_fit_sse(None, None, None, None, 1, 0, None, None, None, None) == -22
    assert L.rap_fmha_schedule(None, 1, 8, 0, None, 0, None, None) == -22


def test_install_eval_patches_reference_modules():
    """rap_b200.install_eval() swaps the metric functions inside the reference's own eval modules (needs /root/reference: build container)."""
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference sources not present")
    import rap_b200
    ns = ref_loader.load_reference_eval()
    try:
        rap_b200.install_eval()
        from rap_b200.rectified_point_flow.eval import metrics as mine
        for n in ("compute_rigidity_rmse", "compute_cd", "compute_overlap_ratio", "compute_transform_errors", "compute_transform_errors_direct"):
            assert getattr(ns.metrics, n) is getattr(mine, n)
            assert not hasattr(ns.evaluator, n) or getattr(ns.evaluator, n) is getattr(mine, n)      # names evaluator.py imported by value
        assert ns.metrics.compute_ecdf.__module__.endswith("eval.metrics") and "rap_b200" not in ns.metrics.compute_ecdf.__module__
    finally:
        ref_loader.unload_reference()


def _fmha_schedule_persistent(cu, n_ctas=148, force=0, heads=8):
    import ctypes as C
    from rap_b200 import _lib
    cu_a = _lib.i32_array(cu)
    n = _lib.lib.rap_fmha_schedule_persistent(cu_a, len(cu) - 1, heads, n_ctas, force, None, 0, None, 0, None, None, None)
    buf, cp = (C.c_int32 * (8 * n))(), (C.c_int32 * (n_ctas + 2))()
    nc, ns, mp = C.c_int(), C.c_int(), C.c_int()
    assert _lib.lib.rap_fmha_schedule_persistent(cu_a, len(cu) - 1, heads, n_ctas, force, buf, n, cp, n_ctas + 2, C.byref(nc), C.byref(ns), C.byref(mp)) == n
    return np.array(buf).reshape(n, 8), np.array(cp)[:nc.value + 1], ns.value, mp.value


@pytest.mark.parametrize("force", [0, 1])
@pytest.mark.parametrize("cu", [[0, 4096, 8192], [0, 8192], [0, 512, 1024], [0, 16384], [0, 3000, 5000, 9000], [0, 100], [0, 130, 700, 701],
                                [0, 4096, 8192, 12288, 16384], [0, 200, 336, 408], list(range(0, 8192 * 8 + 1, 8192))])
def test_fmha_persistent_schedule_tiles_the_work_and_balances_the_ctas(cu, force):
    """Host logic of the default attention launch (one persistent CTA per SM): the items tile (query block, head) x keys exactly once,
    pieces of a cut unit are contiguous in the keys, share a slot and are numbered 0..n-1 in key order, every CTA owns a non-empty contiguous
    item range, whole units are dealt wave by wave (unit k*G + c -> CTA c: all CTAs stay inside one window of K/V streams), only the last
    partial wave is cut, and at the BASELINE shapes no CTA carries more than 3 % above the mean estimated cost (= the kernel's duration)."""
    items, cta_ptr, n_slots, max_pieces = _fmha_schedule_persistent(cu, force=force)
    assert cta_ptr[0] == 0 and cta_ptr[-1] == len(items) and (np.diff(cta_ptr) > 0).all() and len(cta_ptr) - 1 <= 148
    seen = {}
    for q0, s0, s1, head, lo, hi, slot, piece in items.tolist():
        assert s0 <= q0 < s1 and s0 <= lo < hi <= s1 and (lo - s0) % 128 == 0 and (hi == s1 or (hi - s0) % 128 == 0) and (q0 - s0) % 256 == 0
        n_p, idx = piece >> 8, piece & 0xff
        assert (slot < 0) == (n_p == 1) and idx < n_p <= max(max_pieces, 1)
        seen.setdefault((q0, head), []).append((lo, hi, slot, idx, n_p))
    assert len(seen) == 8 * sum(-(-(b - a) // 256) for a, b in zip(cu[:-1], cu[1:]))
    slots = set()
    for (q0, head), pieces in seen.items():
        pieces.sort()
        seg = next((a, b) for a, b in zip(cu[:-1], cu[1:]) if a <= q0 < b)
        assert pieces[0][0] == seg[0] and pieces[-1][1] == seg[1] and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
        assert len({p[2] for p in pieces}) == 1 and [p[3] for p in pieces] == list(range(pieces[0][4]))     # piece order = key order
        if pieces[0][2] >= 0:
            slots.add(pieces[0][2])
    assert slots == set(range(n_slots))
    # the kernel's combine protocol relies on this: a later piece (index >= 1) is the FIRST item of its CTA (it publishes and moves on), the head
    # piece (index 0) of a cut unit is the LAST item of its CTA (it finishes last and combines) -- so a waited-for piece never waits itself
    for a, b in zip(cta_ptr[:-1], cta_ptr[1:]):
        for i in range(a, b):
            slot, idx = items[i, 6], items[i, 7] & 0xff
            if slot >= 0:
                assert i == a if idx > 0 else i == b - 1
    blocks = (items[:, 5] - items[:, 4] + 127) // 128
    cost = blocks * np.where(items[:, 0] + 128 < items[:, 2], 1.0, 0.55) + 2.0
    per = np.array([cost[a:b].sum() for a, b in zip(cta_ptr[:-1], cta_ptr[1:])])
    if len(cu) == 9:      # 8 scenes x 8192: 2048 units = 13 whole waves of 148 + a cut remainder; wave k of CTA c is unit k * 148 + c
        assert len(per) == 148 and per.max() <= 1.03 * per.mean()
        whole = items[cta_ptr[7]:cta_ptr[8]]
        whole = whole[whole[:, 6] < 0]
        order = [(s0 // 8192) * 8 * 32 + head * 32 + (q0 - s0) // 256 for q0, s0, s1, head, *_ in whole.tolist()]
        assert order[:13] == [k * 148 + 7 for k in range(13)]
    if cu in ([0, 4096, 8192], [0, 8192], [0, 16384], [0, 4096, 8192, 12288, 16384]):
        assert len(per) == 148 and per.max() <= 1.03 * per.mean() and n_slots > 0 and 2 <= max_pieces <= 4
    if cu == [0, 512, 1024] and not force:
        assert n_slots == 0 and len(per) == 32            # small launch: whole units, nothing is cut
