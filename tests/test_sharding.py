"""Multi-rank host logic on CPU: world_size-2 gloo process group, pair sharding + pose gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from superpoints_registration_b200.sharding import gather_poses, shard_batch, shard_bounds


def test_shard_bounds_cover_everything_once():
    rng = np.random.default_rng(0)
    for world in (1, 2, 3, 4, 8):
        for n in (1, 2, 7, 8, 33):
            costs = rng.integers(1000, 60000, size=n).tolist()
            b = shard_bounds(costs, world)
            assert len(b) == world and b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(lo <= hi for lo, hi in b)
            if n >= world:
                assert all(hi > lo for lo, hi in b)  # nobody idles when there is enough work


def test_shard_bounds_balance():
    costs = [100] * 64
    b = shard_bounds(costs, 8)
    assert [hi - lo for lo, hi in b] == [8] * 8
    b = shard_bounds([1000, 10, 10, 10, 10, 10, 10, 1000], 2)
    loads = [sum([1000, 10, 10, 10, 10, 10, 10, 1000][lo:hi]) for lo, hi in b]
    assert abs(loads[0] - loads[1]) <= 1000


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_pose(pair_id):
    t = torch.zeros(3, 4)
    t[:, :3] = torch.eye(3)
    t[:, 3] = torch.tensor([pair_id, 2.0 * pair_id, -pair_id], dtype=torch.float32)
    return t


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        sizes = rng.integers(50, 400, size=(n_pairs, 2))
        batch = {"src_xyz": [torch.zeros(int(a), 3) for a, _ in sizes], "tgt_xyz": [torch.zeros(int(b), 3) for _, b in sizes],
                 "pair_id": torch.arange(n_pairs)}
        local, bounds = shard_batch(batch, rank, world)
        lo, hi = bounds[rank]
        assert len(local["src_xyz"]) == hi - lo and local["pair_id"].tolist() == list(range(lo, hi))
        poses = torch.stack([_fake_pose(i) for i in range(lo, hi)]) if hi > lo else torch.zeros((0, 3, 4))
        full = gather_poses(poses, bounds)
        want = torch.stack([_fake_pose(i) for i in range(n_pairs)])
        q.put((rank, bool(torch.equal(full, want)), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [5, 2, 1])
def test_gloo_world2_shard_and_gather(n_pairs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in results:
        assert ok and shape == (n_pairs, 3, 4), (rank, ok, shape)
