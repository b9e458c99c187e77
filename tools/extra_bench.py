"""Secondary measurements for DESIGN.md (run under gpurun): precision tiers, general kernel, output diet, in-kernel ray
generation, point-query (density grid) throughput.  One JSON object on stdout."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from behindthescenes_b200 import synthetic as syn

dev = torch.device("cuda")
wl = syn.cfg2_kitti360_eval(H=192, W=640, K=64, stereo=True, learn_empty=True, lin_out_bias=-4.0, seed=0)
net, wrapped = syn.build_model(wl, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
rays = wl.rays.to(dev)
n_rays = rays.shape[1]
with torch.no_grad():
    net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)


def timed(fn, steps=30, warm=5):
    evs = []
    with torch.no_grad():
        for i in range(warm + steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            if i >= warm:
                evs.append((a, b))
        torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / len(evs)


res = {}
std = dict(want_weights=True, want_alphas=True)
for name, prec in (("tc_exact", 0), ("tc_fast", 1), ("general_fp32", 2)):
    wrapped.renderer.precision = prec
    ms = timed(lambda: wrapped(rays, **std))
    res[name] = {"ms": ms, "rays_per_s": n_rays / ms * 1e3}
wrapped.renderer.precision = 0
ms = timed(lambda: wrapped(rays, want_invalid=False, want_invalid_mass=True))
res["tc_exact_output_diet(rgb,depth,invalid_mass)"] = {"ms": ms, "rays_per_s": n_rays / ms * 1e3}
P, Kc = wl.poses_c2w.to(dev), wl.Ks.to(dev)
ms = timed(lambda: wrapped.render_views(P, Kc, 192, 640, 3.0, 80.0, want_depth_z=True, **std))
res["tc_exact_render_views(in-kernel rays)"] = {"ms": ms, "rays_per_s": n_rays / ms * 1e3}
# density grid query: 256 x 64 x 256 voxels in front of the camera (scripts/inference_setup.py:201-229)
xs = torch.linspace(-9, 9, 256, device=dev); ys = torch.linspace(-0.5, 1.0, 64, device=dev); zs = torch.linspace(3, 21, 256, device=dev)
grid = torch.stack(torch.meshgrid(ys, zs, xs, indexing="ij"), -1)[..., [2, 0, 1]].reshape(1, -1, 3).contiguous()
ms = timed(lambda: net(grid, only_density=True), steps=20)
res["point_query_only_density(256x64x256)"] = {"ms": ms, "points_per_s": grid.shape[1] / ms * 1e3}
ms = timed(lambda: net(grid), steps=20)
res["point_query_full(256x64x256)"] = {"ms": ms, "points_per_s": grid.shape[1] / ms * 1e3}
# BASELINE config 5 shape on a ray subset: d_hidden = 128, K = 256 (random rays against a 192x640 map)
wl5 = syn.cfg5_stress(R=262144, K=256, d_hidden=128)
net5, wrapped5 = syn.build_model(wl5, dev)
with torch.no_grad():
    net5.encode(wl5.images.to(dev), wl5.Ks.to(dev), wl5.poses_c2w.to(dev), ids_encoder=wl5.ids_encoder, ids_render=wl5.ids_render)
rays5 = wl5.rays.to(dev)
for name, prec in (("stress_Hd128_K256_tc_exact", 0), ("stress_Hd128_K256_general", 2)):
    wrapped5.renderer.precision = prec
    ms = timed(lambda: wrapped5(rays5, want_invalid=False), steps=5, warm=2)
    res[name] = {"ms": ms, "rays_per_s": rays5.shape[1] / ms * 1e3, "samples_per_s": rays5.shape[1] * 256 / ms * 1e3}
print(json.dumps(res))
