"""Ray-shard data parallelism: one process per GPU, rays split by contiguous blocks, scene replicated.

Every ray is independent given the per-frame scene, so the only collective of the path is the all-reduce of scalar
loss / metric numerators -- the reference's ``sync_all_reduce("_num_examples:SUM", "_sum:SUM")`` pattern
(``utils/metrics.py:31``).  This replaces the reference's dormant single-process ``DataParallel(dim=1)`` over rays
(``models/common/render/nerf.py:453-457``).
"""
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rays: int, rank: int, world: int, granularity: int = 1) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous block of the ray dimension.  ``granularity`` keeps groups of rays together
    (64 for the 8x8 patches of ``PatchRaySampler``, ``models/bts/model/ray_sampler.py:147-155``); blocks differ by at
    most one group."""
    if n_rays % granularity != 0:
        raise ValueError(f"{n_rays} rays is not a multiple of the granularity {granularity}")
    groups = n_rays // granularity
    base, extra = divmod(groups, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo * granularity, hi * granularity


def shard_rays(rays: torch.Tensor, rank: int, world: int, granularity: int = 1) -> torch.Tensor:
    """rays (n, R, 8) -> this rank's (n, R_r, 8) view."""
    lo, hi = shard_bounds(rays.shape[1], rank, world, granularity)
    return rays[:, lo:hi]


def allreduce_sums(values: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """SUM-all-reduce a dict of scalars in one packed collective (NCCL on GPUs, gloo in the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(values)
    keys = sorted(values)
    buf = torch.stack([values[k].detach().to(torch.float64).reshape(()) for k in keys])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return {k: buf[i].to(values[k].dtype) for i, k in enumerate(keys)}


def gather_ray_outputs(local: torch.Tensor, n_rays: int, granularity: int = 1) -> torch.Tensor:
    """Optional: reassemble a per-ray output (n, R_r, ...) on every rank (uneven blocks allowed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(n_rays, r, world, granularity) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    padded = local.new_zeros((local.shape[0], longest) + tuple(local.shape[2:]))   # all_gather needs equal shapes
    padded[:, : local.shape[1]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:, : hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=1)
