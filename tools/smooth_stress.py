#!/usr/bin/env python
"""Reproducer for the patch_smooth transient: run render work in this process, then the smoothness
forward/backward many times, comparing every launch bit-for-bit with the first one and with the CPU oracle.
usage (GPU box): python tools/smooth_stress.py [iters] [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from behindthescenes_b200 import synthetic as syn  # noqa: E402
from behindthescenes_b200.loss import patch_edge_aware_smoothness  # noqa: E402
from oracle import loss_oracle as lo  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda")
shape = (4, 64, 8, 8)
n, pc, h, w = shape
g = torch.Generator().manual_seed(sum(shape) + 1)
gt = torch.rand((n, pc, h, w, 3), generator=g)
depth = 2 + 90 * torch.rand(shape, generator=g)
depth[0, 0].fill_(7.0)
invalid = torch.rand(shape, generator=g) > 0.7
_, ref = lo.smooth_loss(depth, gt, invalid)

wl = syn.cfg2_kitti360_eval(H=96, W=320, stereo=True, lin_out_bias=-4.0, seed=31, ray_subsample=20001)
net, wrapped = syn.build_model(wl, dev)
with torch.no_grad():
    net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
rays = wl.rays.to(dev)
bad, first, report = 0, None, []
for it in range(iters):
    with torch.no_grad():                      # render work in between (both kernel families, changing allocator state)
        wrapped.renderer.band = bool(it & 1)
        out = wrapped(rays, want_weights=True, want_alphas=True)
        del out
    d = depth.to(dev).requires_grad_(True)
    loss, per_ray = patch_edge_aware_smoothness(d, gt.to(dev), invalid_ray=invalid.to(dev))
    loss.backward()
    got = per_ray.detach().cpu()
    grad = d.grad.cpu()
    if first is None:
        first = (got.clone(), grad.clone())
    dg = (got - first[0]).abs().max().item()
    dgr = (grad - first[1]).abs().max().item()
    do = (got - ref).abs().max().item()
    if dg > 0 or dgr > 0 or do > 1e-5 * max(1.0, ref.max().item()):
        bad += 1
        idx = (got - first[0]).abs().flatten().argmax().item()
        report.append({"iter": it, "vs_first_fwd": dg, "vs_first_grad": dgr, "vs_oracle": do, "worst_idx": idx,
                       "n_diff": int(((got - first[0]).abs() > 0).sum().item())})
res = {"iters": iters, "bad_iterations": bad, "first_vs_oracle": (first[0] - ref).abs().max().item(), "report": report[:20]}
print(json.dumps(res, indent=1))
if len(sys.argv) > 2:
    json.dump(res, open(sys.argv[2], "w"), indent=1)
