"""CPU: the keypoint-selection oracle (oracle/preproc_oracle.py) and the host-side budget rules of the mirror against fixtures produced
by the reference run verbatim (oracle/make_golden.py preproc)."""
import logging
import os

import numpy as np
import pytest

from oracle import preproc_oracle as po
from oracle.make_golden import PREPROC_VOXEL_CASES, preproc_cloud

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "preproc.npz"))


@pytest.mark.parametrize("name", sorted(PREPROC_VOXEL_CASES))
def test_voxel_downsample_oracle_matches_reference(name):
    idx, n, vs, snap = PREPROC_VOXEL_CASES[name]
    pts = preproc_cloud(idx, n, snap)
    assert np.array_equal(po.voxel_down_sample(pts, vs), G[f"voxel_{name}"])
    assert po.voxel_coverage(pts, 4 * vs) == int(G[f"coverage_{name}"])
    assert po.voxel_coverage(pts.astype(np.float64), 4 * vs) == int(G[f"coverage64_{name}"])


def test_budget_rules_match_reference():
    from rap_b200.dataset_process.utils import point_sampling_utils as mine
    logging.disable(logging.WARNING)
    try:
        for trial, row in enumerate(G["alloc_cases"]):
            n, num, mn = (int(v) for v in row[:3])
            pts, cov, want_pc, want_sc = row[3:3 + n], row[11:11 + n], row[19:19 + n], row[27:27 + n]
            np.random.seed(trial)
            assert np.array_equal(mine._allocate_by_point_count_logic(pts.copy(), num, mn), want_pc), trial
            np.random.seed(trial)
            assert np.array_equal(mine._allocate_by_spatial_coverage_logic(cov.copy(), pts.copy(), num, mn), want_sc), trial
    finally:
        logging.disable(logging.NOTSET)
    parts = [preproc_cloud(6, 4000), preproc_cloud(7, 900), np.zeros((0, 3), np.float32)]
    assert po.adaptive_sample_counts(parts, 2.0, 0.5, 100, 1500) == list(G["adaptive_counts"])
    assert list(mine.allocate_fps_points(parts[:2], "voxel_adaptive", 0, 100, 2.0, 0.5, total_sample_points=[5000, 300])) == [4000, 300]
    with pytest.raises(ValueError):
        mine.allocate_fps_points([10, 20], "voxel_adaptive", 0, 1, 1.0, 1.0, total_sample_points=[1, 2])
    with pytest.raises(ValueError):
        mine.allocate_fps_points([10, 20], "nope", 0, 1, 1.0, 1.0)


def test_fps_oracle_properties():
    """The restated pytorch3d semantics: start index first, no repeats, -1 padding, first-index ties, non-increasing pick distances."""
    rng = np.random.default_rng(3)
    pts = rng.standard_normal((3, 400, 3)).astype(np.float32)
    pts[2, :50] = pts[2, 0]                                         # duplicates: distance-0 ties
    lengths, K = np.array([400, 250, 60]), np.array([64, 300, 60])
    start = po.fps_start_indices(lengths, seed=42)
    assert np.array_equal(start, po.fps_start_indices(lengths, seed=42)) and (start < lengths).all()
    idx = po.farthest_point_sampling(pts, lengths, K, start)
    assert idx.shape == (3, 300)
    for n in range(3):
        k_n = min(K[n], lengths[n])
        sel = idx[n, :k_n]
        assert sel[0] == start[n] and len(set(sel.tolist())) == k_n and (sel < lengths[n]).all() and (idx[n, k_n:] == -1).all()
    d = np.linalg.norm(pts[0, idx[0, :64], None] - pts[0, None, idx[0, :64]], axis=-1)
    radii = [d[k, :k].min() for k in range(1, 64)]
    assert all(a >= b - 1e-6 for a, b in zip(radii, radii[1:]))
    assert set(idx[2, :60].tolist()) == set(range(60))             # every point picked once even though 50 coincide


def test_outlier_oracle_flags_isolated_points():
    rng = np.random.default_rng(5)
    pts = rng.standard_normal((2000, 3)).astype(np.float32)
    pts[:5] += 40.0
    keep = po.statistical_outlier_mask(pts, 20, 2.0)
    assert not keep[:5].any() and keep.sum() > 1800
