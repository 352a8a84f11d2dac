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
