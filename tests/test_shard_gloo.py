"""CPU: the N>1 path of SURVEY.md 8-e with a world_size-2 gloo group -- contiguous block partition of the
reference views and the single all-gather of per-view depth maps (wild_deep_mvs_b200/shard.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wild_deep_mvs_b200 import shard


def test_block_partition_is_contiguous_and_balanced():
    for n in (1, 2, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            blocks = [shard.block_partition(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = shard.block_partition(n_items, world, rank)
        # "depth map" of global sample i is the constant image i + 1
        local = torch.stack([torch.full((4, 6), float(i + 1)) for i in range(a, b)]) if b > a else torch.zeros(0, 4, 6)
        out = shard.gather_depth_maps(local, n_items)
        q.put((rank, out.shape[0], out[:, 0, 0].tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [4, 5])   # equal blocks (one all_gather_into_tensor) and ragged blocks (padded)
def test_gather_depth_maps_world2_gloo(n_items):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n, vals in results:
        assert n == n_items
        assert vals == [float(i + 1) for i in range(n_items)], (rank, vals)   # ordered by global sample index on every rank


def test_gather_is_identity_without_a_process_group():
    x = torch.arange(6.0).view(2, 3)
    assert shard.gather_depth_maps(x) is x


def _gather_worker(rank, world, port, n_local, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = shard.DepthGather(n_local, (4, 6), "cpu")
        base = g.all.data_ptr()
        res = []
        for step in range(3):    # the buffer is reused step after step: no allocation, the producer writes the slice in place
            for j in range(n_local):
                g.local[j].fill_(100.0 * step + rank * n_local + j + 1)    # "K3 writes its map into the send slice"
            out = g.all_gather()
            assert out.data_ptr() == base and g.local.data_ptr() == base + rank * n_local * 4 * 6 * 4
            res.append(out[:, 0, 0].tolist())
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_local", [1, 4])
def test_depth_gather_in_place_world2_gloo(n_local):
    """shard.DepthGather: the preallocated buffer + in-place all-gather that bench.py captures into the step's CUDA graph."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, n_local, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        for step, vals in enumerate(res):
            assert vals == [100.0 * step + i + 1 for i in range(world * n_local)], (rank, step, vals)


def test_depth_gather_single_process():
    g = shard.DepthGather(2, (3, 5), "cpu")
    assert g.world == 1 and g.local.shape == (2, 3, 5) and g.local.data_ptr() == g.all.data_ptr()
    g.local.fill_(7.0)
    assert g.all_gather() is g.all and float(g.all.sum()) == 7.0 * 30


def _consumer_scene(N=2, h=12, w=16):
    """N pinhole views of a slanted plane: per-view depth maps (what each rank would compute) and 4x4 projections."""
    import numpy as np
    proj = np.zeros((1, N, 4, 4), np.float32)
    maps = np.zeros((N, h, w), np.float32)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    n, c = np.array([0.1, -0.05, 1.0]), 500.0
    for v in range(N):
        K = np.array([[120.0, 0, w / 2.0], [0, 118.0, h / 2.0], [0, 0, 1]])
        a = 0.03 * v
        R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
        t = np.array([[-15.0 * v], [1.0 * v], [0.0]])
        P = np.eye(4)
        P[:3, :3], P[:3, 3:] = K @ R, K @ t
        proj[0, v] = P
        rays = np.stack([xs, ys, np.ones_like(xs)], -1) @ np.linalg.inv(K).T @ R
        o = -R.T @ t.reshape(3)
        maps[v] = (c - n @ o) / (rays @ n)
    maps[1, 3:6, 4:9] *= 1.3       # view 1's estimate is wrong in a block: the re-projection test must fail there
    return proj, maps


def _consumer_worker(rank, world, port, q):
    """models/trainer.py:101,246-274 of the reference on two ranks: rank r's reference view is view r, the ONE all-gather
    brings every rank's depth map, and each rank masks ITS map against the gathered ones (the K9 operator; here its CPU
    oracle, the collective is what is under test)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.filter import gathered_masks
        proj, maps = _consumer_scene(world)
        g = shard.DepthGather(1, maps.shape[1:], "cpu")
        g.local[0].copy_(torch.from_numpy(maps[rank]))                 # "K3 writes its map into the send slice"
        gathered = g.all_gather().numpy()                              # [world, h, w] on every rank
        out = gathered_masks(maps[rank][None], gathered[None], proj, rank, 0.05)
        q.put((rank, gathered.tolist(), out["masks"].tolist()))
    finally:
        dist.destroy_process_group()


def test_gathered_maps_feed_the_consumer_on_every_rank_world2_gloo():
    import numpy as np
    from oracle.filter import gathered_masks
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_consumer_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {r: (np.asarray(gm, np.float32), np.asarray(m)) for r, gm, m in (q.get(timeout=120) for _ in range(world))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    proj, maps = _consumer_scene(world)
    for rank in range(world):
        gathered, masks = results[rank]
        assert np.array_equal(gathered, maps)                          # every rank holds every view's map, in view order
        want = gathered_masks(maps[rank][None], maps[None], proj, rank, 0.05)["masks"]
        assert np.array_equal(masks, want)
    # the corrupted block of view 1 is rejected from both sides, the rest of the overlap is kept
    assert not results[1][1][0, 0, 3:6, 4:9].any() and results[0][1].mean() > 0.3
