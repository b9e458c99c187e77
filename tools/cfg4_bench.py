#!/usr/bin/env python
"""cfg4 (RealEstate10K shape) render: tensor-core gather kernel vs the general CUDA-core kernel.  usage: python tools/cfg4_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from behindthescenes_b200 import _lib, synthetic as syn
from tools.bench_workloads import timed_steps
dev = torch.device("cuda", 0)
res = {}
for K in (128, 48):
    wl = syn.cfg4_re10k_nvs(H=256, W=384, K=K, lin_out_bias=-4.0, seed=14)
    net, wrapped = syn.build_model(wl, dev)
    with torch.no_grad():
        net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        rays = wl.rays.to(dev)
        for prec, name in ((0, "tc_exact"), (1, "tc_fast"), (2, "general")):
            wrapped.renderer.precision = prec
            ms = timed_steps(lambda: wrapped(rays, want_weights=True, want_alphas=True), dev, 20, 3)
            res[f"K{K}_{name}"] = {"ms": ms, "Mrays_s": rays.shape[1] / ms / 1e3, "kernel": _lib.load().bts_last_kernel().decode()}
print(json.dumps(res, indent=1))
