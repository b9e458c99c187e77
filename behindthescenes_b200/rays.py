"""Ray-shard data parallelism: one process per GPU, rays split by contiguous blocks, scene replicated.

Every ray is independent given the per-frame scene, so the only collective of the path is the all-reduce of scalar
loss / metric numerators -- the reference's ``sync_all_reduce("_num_examples:SUM", "_sum:SUM")`` pattern
(``utils/metrics.py:31``).  This replaces the reference's dormant single-process ``DataParallel(dim=1)`` over rays
(``models/common/render/nerf.py:453-457``).
"""
from typing import Dict, Iterable, Tuple

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


def allreduce_gradients(parameters: Iterable[torch.Tensor], n_local_rays: int) -> int:
    """Training under a ray shard (BASELINE config 3): every rank back-propagated the MEAN loss of its own rays; turn the
    local gradients into the gradient of the mean over ALL rays with ONE packed all-reduce (NCCL over NVLink; gloo in the
    CPU tests): g = sum_r (n_r / n_total) g_r.  Parameters without a gradient on some rank contribute zeros.  Returns the
    global ray count.  This is the per-process form of what ``DataParallel(dim=1)`` (``nerf.py:453-457``) would have
    summed on device 0; the reference trains single-GPU (``utils/base_trainer.py:297``)."""
    params = [p for p in parameters if p.requires_grad]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1 or not params:
        return int(n_local_rays)
    dev = params[0].device
    sizes = [p.numel() for p in params]
    flat = torch.zeros(sum(sizes) + 1, device=dev, dtype=torch.float32)
    off = 0
    for p, n in zip(params, sizes):
        if p.grad is not None:
            flat[off:off + n] = p.grad.reshape(-1).to(torch.float32)
        off += n
    flat[:-1] *= float(n_local_rays)
    flat[-1] = float(n_local_rays)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    n_total = flat[-1].clone()
    flat[:-1] /= n_total
    off = 0
    for p, n in zip(params, sizes):
        g = flat[off:off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return int(round(n_total.item()))


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
