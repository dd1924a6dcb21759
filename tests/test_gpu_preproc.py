"""Keypoint selection on the B200 through the reference-facing mirrors (dataset_process.utils.*, SampleProcessor): index results
bit-exact against the reference's own outputs (golden) and the CPU oracle, plus size-independent properties at full size."""
import os
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import preproc_oracle as po  # noqa: E402
from oracle.make_golden import PREPROC_VOXEL_CASES, preproc_cloud  # noqa: E402
from rap_b200.synth import make_raw_cloud, make_spinnet_state_dict  # noqa: E402


def _sync_check():
    from rap_b200 import _lib
    torch.cuda.synchronize()
    assert _lib.device_error() == 0


@pytest.mark.parametrize("name", sorted(PREPROC_VOXEL_CASES))
def test_voxel_downsample_matches_reference_golden(golden_dir, name):
    from rap_b200.dataset_process.utils import dataset_utils as du, calculate_voxel_coverage
    g = np.load(os.path.join(golden_dir, "preproc.npz"))
    idx, n, vs, snap = PREPROC_VOXEL_CASES[name]
    pts = preproc_cloud(idx, n, snap)
    got = du.voxel_down_sample_torch(torch.from_numpy(pts).cuda(), vs)
    _sync_check()
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), g[f"voxel_{name}"])
    down, _ = du.downsample_points(pts, None, method="voxel", voxel_size=vs)
    assert np.array_equal(down, pts[g[f"voxel_{name}"]])
    assert calculate_voxel_coverage(pts, 4 * vs) == int(g[f"coverage_{name}"])
    assert calculate_voxel_coverage(pts.astype(np.float64), 4 * vs) == int(g[f"coverage64_{name}"])


def test_voxel_downsample_large_vs_oracle_and_properties():
    from rap_b200.dataset_process.utils import dataset_utils as du
    rng = np.random.default_rng(1)
    pts = (rng.standard_normal((1_000_000, 3)) * np.array([40, 40, 6])).astype(np.float32)
    t0 = time.perf_counter()
    got = du.voxel_down_sample_torch(torch.from_numpy(pts).cuda(), 0.5).cpu().numpy()
    dt = time.perf_counter() - t0
    want = po.voxel_down_sample(pts, 0.5)
    print(f"[voxel 1M pts] {len(got)} voxels, {dt*1e3:.1f} ms incl. H2D/D2H")
    assert np.array_equal(got, want)
    cells = np.floor(pts[got] / np.float32(0.5)).astype(np.int64)
    assert len(np.unique(cells, axis=0)) >= len(got) - 8            # one pick per voxel (the reference's id aliasing merges a few)


FPS_CASES = {
    # name: (lengths, K)  -> cluster size / points-per-thread path
    "one_cta": ([3000, 4096, 1, 777], [256, 300, 1, 777]),                 # C=1, PPT=8; K == length; single point
    "cluster2_ragged": ([8000, 5000, 6500], [200, 64, 128]),               # C=2
    "cluster8": ([30000, 20000], [160, 96]),                               # C=8, PPT=8
    "regs20": ([60000, 33000], [96, 64]),                                  # 20 points per thread in registers, C=8
    "smem36": ([90000, 82000], [80, 48]),                                  # 36 points per thread, coordinates in shared memory, C=8
}


@pytest.mark.parametrize("name", list(FPS_CASES))
def test_fps_matches_oracle(name):
    from rap_b200.dataset_process.utils.point_sampling_utils import sample_farthest_points
    lengths, K = (np.array(v) for v in FPS_CASES[name])
    P = int(lengths.max())
    rng = np.random.default_rng(len(name))
    pts = np.zeros((len(lengths), P, 3), np.float32)
    for n, L in enumerate(lengths):
        pts[n, :L] = (rng.standard_normal((L, 3)) * (1 + n)).astype(np.float32)
    torch.manual_seed(7)
    sampled, idx = sample_farthest_points(torch.from_numpy(pts).cuda(), lengths=torch.from_numpy(lengths).cuda(), K=torch.from_numpy(K),
                                          random_start_point=True)
    _sync_check()
    start = po.fps_start_indices(lengths, seed=7)
    want = po.farthest_point_sampling(pts, lengths, K, start)
    got = idx.cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"first mismatch at {np.argwhere(got != want)[:3]}"
    for n in range(len(lengths)):
        k_n = int(min(K[n], lengths[n]))
        assert np.array_equal(sampled[n, :k_n].cpu().numpy(), pts[n, want[n, :k_n]]) and not sampled[n, k_n:].any()


def test_fps_duplicates_and_fixed_start():
    from rap_b200.dataset_process.utils.point_sampling_utils import sample_farthest_points
    rng = np.random.default_rng(9)
    pts = rng.standard_normal((2, 600, 3)).astype(np.float32)
    pts[0, 100:400] = pts[0, 5]                                     # 300 coincident points: distance-0 ties go to the lowest free index
    pts[1] = np.round(pts[1])                                       # lattice: many exact distance ties
    lengths, K = np.array([600, 600]), np.array([600, 200])
    _, idx = sample_farthest_points(torch.from_numpy(pts).cuda(), lengths=torch.from_numpy(lengths), K=torch.from_numpy(K))
    want = po.farthest_point_sampling(pts, lengths, K, np.zeros(2, np.int64))
    assert np.array_equal(idx.cpu().numpy(), want)
    assert sorted(idx[0].tolist()) == list(range(600))


def test_fps_full_size_properties():
    """Reference-scale job (pre-FPS cap 20 x 5000 points, 5000 keypoints, extract_sample_features.py:417-437): no repeats, start first,
    non-increasing pick radii, and the first 64 picks against the oracle."""
    from rap_b200.dataset_process.utils.point_sampling_utils import apply_batched_fps
    rng = np.random.default_rng(2)
    lengths = np.array([100_000, 41_000, 73_500])
    K = np.array([5000, 2000, 3000])
    pts = np.zeros((3, 100_000, 3), np.float32)
    for n, L in enumerate(lengths):
        pts[n, :L] = (rng.standard_normal((L, 3)) * np.array([30, 30, 4])).astype(np.float32)
    batch = torch.from_numpy(pts)
    apply_batched_fps(batch[:, :1000], torch.tensor([1000] * 3), torch.tensor([10] * 3), 42, torch.device("cuda"))   # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sampled, idx = apply_batched_fps(batch.cuda(), torch.from_numpy(lengths), torch.from_numpy(K), 42, torch.device("cuda"))
    _sync_check()
    print(f"[fps 3 clouds <=100k pts, K<=5000] {1e3 * (time.perf_counter() - t0):.1f} ms")
    start = po.fps_start_indices(lengths, seed=42)
    head = po.farthest_point_sampling(pts, lengths, np.array([64, 64, 64]), start)
    for n in range(3):
        sel = idx[n, :K[n]].cpu().numpy()
        assert sel[0] == start[n] and len(np.unique(sel)) == K[n] and sel.max() < lengths[n] and (idx[n, K[n]:] == -1).all()
        assert np.array_equal(sel[:64], head[n])
        assert np.array_equal(sampled[n].cpu().numpy(), pts[n, sel])
        chosen = pts[n, sel[:400]].astype(np.float64)
        d = np.linalg.norm(chosen[:, None] - chosen[None], axis=-1)
        radii = np.array([d[k, :k].min() for k in range(1, 400)])
        assert (radii[:-1] >= radii[1:] - 1e-5).all()


def test_statistical_outliers_match_oracle():
    from rap_b200.dataset_process import remove_statistical_outlier
    rng = np.random.default_rng(4)
    pts = (rng.standard_normal((12_000, 3)) * np.array([20, 20, 3])).astype(np.float32)
    pts[:40] += rng.standard_normal((40, 3)).astype(np.float32) * 60
    for k, ratio in ((20, 2.0), (8, 1.0), (32, 3.0)):
        got = remove_statistical_outlier(pts, k, ratio)
        want = np.nonzero(po.statistical_outlier_mask(pts, k, ratio))[0]
        _sync_check()
        assert np.array_equal(got, want), (k, ratio, len(got), len(want))
    assert len(got) < len(pts)


def test_sample_processor_pipeline_against_composed_oracle():
    """SampleProcessor.process_sample = outlier removal -> budget -> FPS -> descriptors; every index stage must equal the oracle
    composition, and the descriptors must equal a direct extractor call on the same keypoints."""
    from rap_b200.dataset_process import FeatureExtractor, SampleProcessor
    fx = FeatureExtractor(des_r=2.0, state_dict=make_spinnet_state_dict())
    parts = [preproc_cloud(8, 9000), preproc_cloud(9, 2500)]
    sp = SampleProcessor(fx, remove_outliers=True, outlier_nb_neighbors=20, outlier_std_ratio=2.0, min_points_per_part=50,
                         max_points_per_part=400, global_seed=13, allocation_method="voxel_adaptive", voxel_size=1.0, voxel_ratio=0.5)
    np.random.seed(0)
    res = sp.process_sample(parts)
    _sync_check()
    assert len(res) == 2
    # oracle composition
    np.random.seed(0)
    fps_parts = [p[po.statistical_outlier_mask(p, 20, 2.0)] for p in parts]
    cap = 20 * 400
    fps_parts = [p[np.random.choice(len(p), cap, replace=False)] if len(p) > cap else p for p in fps_parts]
    counts = po.adaptive_sample_counts(fps_parts, 1.0, 0.5, 50, 400)
    lengths = np.array([len(p) for p in fps_parts])
    batch = np.zeros((2, lengths.max(), 3), np.float32)
    for i, p in enumerate(fps_parts):
        batch[i, :len(p)] = p
    idx = po.farthest_point_sampling(batch, lengths, np.array(counts), po.fps_start_indices(lengths, 13))
    for i in range(2):
        want_kp = fps_parts[i][idx[i, :counts[i]]]
        assert res[i]["sampled_points"].shape == (counts[i], 3) and np.array_equal(res[i]["sampled_points"], want_kp)
        assert res[i]["features"].shape == (counts[i], 32) and res[i]["sampled_normals"] is None
        np.testing.assert_allclose(np.linalg.norm(res[i]["features"], axis=1), 1.0, atol=1e-4)
    # skip_point_sampling path: all points are keypoints
    sp2 = SampleProcessor(fx, skip_point_sampling=True)
    out2 = sp2.process_sample([parts[1][:300]])
    assert out2[0]["features"].shape == (300, 32) and np.array_equal(out2[0]["sampled_points"], parts[1][:300])
