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


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        lin = torch.nn.Linear(8, 3)
        unused = torch.nn.Parameter(torch.zeros(2))                 # a parameter only rank 0 touches
        g = torch.Generator().manual_seed(0)
        rays = torch.rand((2, 101, 8), generator=g)
        mine = R.shard_rays(rays, rank, world)                       # uneven blocks: 51 / 50 rays
        loss = lin(mine).square().mean() + (unused.sum() if rank == 0 else 0.0)
        loss.backward()
        n_total = R.allreduce_gradients([lin.weight, lin.bias, unused], mine.shape[0] * mine.shape[1])
        q.put((rank, n_total, lin.weight.grad.clone(), lin.bias.grad.clone(), unused.grad.clone()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch_gradient():
    """Sharded mean-loss gradients, ray-count weighted and all-reduced in one packed collective, equal the gradient of the
    mean over all rays (uneven shards, a parameter without gradient on one rank)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    lin = torch.nn.Linear(8, 3)
    rays = torch.rand((2, 101, 8), generator=torch.Generator().manual_seed(0))
    lin(rays).square().mean().backward()
    for _, n_total, gw, gb, gu_ in res:
        assert n_total == 202
        assert torch.allclose(gw, lin.weight.grad, atol=1e-6) and torch.allclose(gb, lin.bias.grad, atol=1e-6)
        assert torch.allclose(gu_, torch.full((2,), 102 / 202), atol=1e-6)      # rank 0's share of the ray count
