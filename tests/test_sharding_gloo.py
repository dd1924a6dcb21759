"""World-size-2 gloo tests of the multi-GPU host logic behind `Registrar.register_sharded`: scene sharding covers every scene exactly
once (even and uneven counts), micro-batching respects the point budget, scene slices re-base their offsets, and the pose all-gather
returns the scene-ordered concatenation on every rank -- also when the shards are uneven."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rap_b200.sharding import gather_poses, micro_batches, shard_counts, shard_range, slice_scenes
from rap_b200.synth import make_batch


def _worker(rank, world, port, n_scenes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine, counts = list(shard_range(n_scenes, rank, world)), shard_counts(n_scenes, world)
    # every rank "registers" its own scenes: pose of scene s = (s+1) * I, translation = s
    R = torch.stack([torch.eye(3) * (s + 1) for s in mine])[:, None].repeat(1, 2, 1, 1) if mine else torch.zeros(0, 2, 3, 3)
    t = torch.stack([torch.full((3,), float(s)) for s in mine])[:, None].repeat(1, 2, 1) if mine else torch.zeros(0, 2, 3)
    poses = gather_poses(R, t, counts=counts)
    ok = poses.shape == (n_scenes, 2, 12)
    for i in range(n_scenes):
        ok &= bool(poses[i, 0, 0] == i + 1) and bool(poses[i, 1, 9] == i)
    out.put((rank, mine, ok))
    dist.destroy_process_group()


def _run(n_scenes, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_scenes) % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_scenes, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(30) for p in procs]
    assert sorted(s for r in res for s in r[1]) == list(range(n_scenes))
    assert all(r[2] for r in res)


def test_scene_sharding_and_pose_gather_world2():
    _run(8)


def test_uneven_shards_pose_gather_world2():
    """7 scenes on 2 ranks (4 + 3) and 1 scene on 2 ranks (1 + 0): the short shard is padded for the collective and stripped after."""
    _run(7)
    _run(1)


def test_shard_range_covers_uneven_counts():
    for n, w in [(7, 2), (256, 8), (3, 4), (64, 8), (0, 2)]:
        got = [s for r in range(w) for s in shard_range(n, r, w)]
        assert got == list(range(n)) and sum(shard_counts(n, w)) == n


def test_slice_scenes_rebases_offsets():
    full = make_batch(5, 3, [40, 24, 16], max_parts=4)
    sub = slice_scenes(full, 1, 4)
    ref = make_batch(3, 3, [40, 24, 16], max_parts=4, first_scene=1)
    assert sub.cu_seqlens_batch.tolist() == ref.cu_seqlens_batch.tolist() and sub.cu_seqlens_batch.dtype == torch.int32
    assert sub.cu_seqlens_part.tolist() == ref.cu_seqlens_part.tolist()
    assert torch.equal(sub.points_per_part, ref.points_per_part)
    for k in ("cond", "features", "scales", "anchor_indices", "x_1"):
        assert torch.equal(getattr(sub, k), getattr(ref, k)), k


def test_micro_batches_respect_the_point_budget():
    cu = torch.tensor([0, 100, 250, 300, 900, 950, 1000])
    assert micro_batches(cu, 0, 6, None) == [(0, 6)]
    mb = micro_batches(cu, 0, 6, 300)
    assert mb == [(0, 3), (3, 4), (4, 6)]                  # scene 3 (600 points) exceeds the budget on its own: a micro-batch of one
    assert [s for a, b in mb for s in range(a, b)] == list(range(6))
    assert micro_batches(cu, 2, 5, 10_000) == [(2, 5)] and micro_batches(cu, 3, 3, 100) == []
    for a, b in micro_batches(cu, 0, 6, 300):
        assert b - a == 1 or int(cu[b] - cu[a]) <= 300
