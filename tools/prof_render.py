#!/usr/bin/env python
"""Render the bench frame (or its encoder-view / stereo half) a few times -- the command ncu wraps.
usage: python tools/prof_render.py [all|centre|stereo] [band|gather] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from behindthescenes_b200 import synthetic as syn  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
band = (sys.argv[2] if len(sys.argv) > 2 else "band") == "band"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 4
dev = torch.device("cuda", 0)
wl = syn.cfg2_kitti360_eval(H=192, W=640, K=64, stereo=True, learn_empty=True, lin_out_bias=-4.0, seed=0)
net, wrapped = syn.build_model(wl, dev)
wrapped.renderer.band = band
with torch.no_grad():
    net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    rays = wl.rays.to(dev)
    n = rays.shape[1]
    r = {"all": rays, "centre": rays[:, : n // 2].contiguous(), "stereo": rays[:, n // 2:].contiguous()}[which]
    for _ in range(iters):
        wrapped(r, want_weights=True, want_alphas=True)
    torch.cuda.synchronize()
print("done", which, band)
