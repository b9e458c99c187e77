"""CPU, world_size 2 over gloo: the host logic of the ray-shard data-parallel path (no CUDA)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from behindthescenes_b200 import rays as R


def test_shard_bounds_cover_and_balance():
    for n, world, g in ((122880, 8, 1), (2048 * 16, 8, 64), (1000, 3, 1), (7 * 64, 4, 64), (5, 8, 1)):
        spans = [R.shard_bounds(n, r, world, g) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= g and all(s % g == 0 for s in sizes)
    with pytest.raises(ValueError):
        R.shard_bounds(100, 0, 2, 64)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        rays = torch.rand((2, 101, 8), generator=g)
        mine = R.shard_rays(rays, rank, world)
        depth_local = mine[..., :3].sum(-1) * 2.0                      # stand-in for a per-ray output
        sums = R.allreduce_sums({"sum": depth_local.sum(), "count": torch.tensor(float(depth_local.numel()))})
        full = R.gather_ray_outputs(depth_local, rays.shape[1])
        ok = torch.allclose(full, rays[..., :3].sum(-1) * 2.0)
        q.put((rank, sums["sum"].item(), sums["count"].item(), bool(ok), tuple(mine.shape)))
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_allreduce_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(0)
    rays = torch.rand((2, 101, 8), generator=g)
    want = (rays[..., :3].sum(-1) * 2.0).sum().item()
    assert res[0][4] == (2, 51, 8) and res[1][4] == (2, 50, 8)
    for _, s, c, ok, _ in res:
        assert abs(s - want) < 1e-3 and c == 202.0 and ok
