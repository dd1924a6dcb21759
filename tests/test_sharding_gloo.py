"""World-size-2 gloo test of the multi-GPU host logic: scene sharding covers every scene exactly once and the pose
all-gather returns the rank-ordered concatenation on every rank."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rap_b200.sharding import gather_poses, shard_range


def _worker(rank, world, port, n_scenes, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = list(shard_range(n_scenes, rank, world))
    # identical-size shards (weak scaling): pad to the common length like bench.py's fixed pairs-per-GPU
    B = n_scenes // world
    R = torch.stack([torch.eye(3) * (s + 1) for s in mine[:B]])[:, None].repeat(1, 2, 1, 1)
    t = torch.stack([torch.full((3,), float(s)) for s in mine[:B]])[:, None].repeat(1, 2, 1)
    poses = gather_poses(R, t)
    ok = poses.shape == (world * B, 2, 12)
    for i in range(world * B):
        ok &= bool(poses[i, 0, 0] == i + 1) and bool(poses[i, 1, 9] == i)
    out.put((rank, mine, ok))
    dist.destroy_process_group()


def test_scene_sharding_and_pose_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 8, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(30) for p in procs]
    assert sorted(res[0][1] + res[1][1]) == list(range(8))
    assert all(r[2] for r in res)


def test_shard_range_covers_uneven_counts():
    for n, w in [(7, 2), (256, 8), (3, 4), (64, 8)]:
        got = [s for r in range(w) for s in shard_range(n, r, w)]
        assert got == list(range(n))
