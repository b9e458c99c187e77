"""Secondary bench workloads (the BASELINE.json configs other than the headline one), reported by ``bench.py`` as extra
keys of its JSON line: strong scaling of one cfg2 frame and of the cfg5 ray batch, the cfg3 training step, the cfg4
(RealEstate10K-shaped) render.  Every function returns a plain dict; timings are CUDA events on the launching stream,
max over ranks where a process group is up.  ``oracle`` is used only as the checker / eager-reference baseline.
"""
import math
import time

import torch
import torch.distributed as dist

from behindthescenes_b200 import rays as R
from behindthescenes_b200 import synthetic as syn

# algorithmic dense FLOPs per sample = 2 (C + P) Hd + 4 nb Hd^2 + 2 Hd
FLOP_SAMPLE = {"kitti": 2 * (64 + 39) * 64 + 2 * 64, "re10k": 2 * (32 + 39) * 32 + 4 * 32 * 32 + 2 * 32,
               "stress": 2 * (64 + 39) * 128 + 2 * 128}


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if _world() > 1 else 0


def _max_over_ranks(ms, dev):
    if _world() == 1:
        return ms
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sync(dev):
    if _world() > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def timed_steps(fn, dev, steps, warm, flush=None):
    """ms per call of ``fn`` (events around each call on the current stream, optional L2 flush outside the pair)."""
    for _ in range(warm):
        fn()
    _sync(dev)
    evs = []
    for _ in range(steps):
        if flush is not None:
            flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    _sync(dev)
    return _max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / steps, dev)


def graphed(fn, dev):
    """Capture ``fn`` (fold + pack + render of fixed buffers) in a CUDA graph so that a small per-GPU ray block is not
    bound by launch / Python overhead.  Returns (replay callable, captured?)."""
    try:
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g.replay, True
    except Exception:      # capture is an optimisation of the bench loop, never a requirement
        torch.cuda.synchronize(dev)
        return fn, False


def roofline(rays_per_s, flop_per_ray, bytes_per_ray, peaks):
    tf, gb = rays_per_s * flop_per_ray / 1e12, rays_per_s * bytes_per_ray / 1e9
    ft, fh = tf / peaks["tflops"], gb / peaks["hbm_gbs"]
    return {"bound": "tensor" if ft >= fh else "hbm", "achieved": tf if ft >= fh else gb, "peak": peaks["tflops"] if ft >= fh else peaks["hbm_gbs"],
            "unit": "TFLOP/s" if ft >= fh else "GB/s", "frac": max(ft, fh), "tensor_frac": ft, "hbm_frac": fh,
            "algorithmic_dense_flop_per_ray": flop_per_ray, "algorithmic_bytes_per_ray": bytes_per_ray, "traffic": None}


# ---------------------------------------------------------------------------------------------------------------------
# strong scaling: ONE cfg2 stereo frame (245 760 rays) split over the ranks (nerf.py:453-457 splits one batch)
# ---------------------------------------------------------------------------------------------------------------------

def strong_cfg2(net, wrapped, frame_rays_dev, dev, steps, flush):
    world, rank = _world(), _rank()
    n_frame = frame_rays_dev.shape[1]
    part = R.shard_rays(frame_rays_dev, rank, world, granularity=8).contiguous()
    wants = dict(want_weights=True, want_alphas=True)

    def step():
        net.invalidate_scene()         # a new frame: fold + packing run again on every rank (scene replicated)
        net._scene_struct()
        return wrapped(part, **wants)

    with torch.no_grad():
        run, captured = graphed(step, dev)
        ms = timed_steps(run, dev, steps, 3, flush)
    return {"what": "one cfg2 stereo frame (245760 rays, K=64) split into contiguous ray blocks, scene replicated; step = fold + pack + render",
            "rays_total": n_frame, "rays_per_gpu": part.shape[1], "ms_per_frame": ms, "rays_per_s": n_frame / ms * 1e3,
            "cuda_graph": captured}


# ---------------------------------------------------------------------------------------------------------------------
# cfg5 stress: 4096 x 4096 rays, K = 256, d_hidden = 128 (tensor-core gather kernel, 128-wide variant)
# ---------------------------------------------------------------------------------------------------------------------

def cfg5_stress(dev, peaks, total_rays, steps=2):
    world, rank = _world(), _rank()
    lo, hi = R.shard_bounds(total_rays, rank, world, granularity=8)
    wl = syn.cfg5_stress(R=8, K=256, d_hidden=128)                       # scene / model of the stress config
    net, wrapped = syn.build_model(wl, dev)
    with torch.no_grad():
        net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    # this rank's rays, generated on the device from the global ray index (same construction as synthetic.cfg5_stress:
    # random source-frustum pixels, origins within +-1 m, directions rotated by <= 5 deg)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    n = hi - lo
    o = torch.rand((n, 3), generator=g, device=dev) * 2 - 1
    px = torch.rand((n, 2), generator=g, device=dev) * 2 - 1
    Kt = torch.tensor(syn.KITTI_RAW_K, device=dev)
    d = torch.stack(((px[:, 0] - Kt[0, 2]) / Kt[0, 0], (px[:, 1] - Kt[1, 2]) / Kt[1, 1], torch.ones(n, device=dev)), -1)
    ang = (torch.rand((n, 2), generator=g, device=dev) * 2 - 1) * math.radians(5.0)
    cy, sy, cp, sp = torch.cos(ang[:, 0]), torch.sin(ang[:, 0]), torch.cos(ang[:, 1]), torch.sin(ang[:, 1])
    dx = cy * d[:, 0] + sy * d[:, 2]
    dz = -sy * d[:, 0] + cy * d[:, 2]
    dy = cp * d[:, 1] - sp * dz
    dz = sp * d[:, 1] + cp * dz
    d = torch.stack((dx, dy, dz), -1)
    d = d / d.norm(dim=-1, keepdim=True)
    nf = torch.tensor([3.0, 80.0], device=dev).view(1, 2).expand(n, 2)
    rays = torch.cat((o, d, nf), -1).view(1, n, 8).contiguous()
    del o, px, d, ang, dx, dy, dz
    with torch.no_grad():
        ms = timed_steps(lambda: wrapped(rays, want_invalid=False), dev, steps, 1)
    from behindthescenes_b200 import _lib
    kern = _lib.load().bts_last_kernel().decode()
    rps = total_rays / ms * 1e3
    flop_ray, bytes_ray = 256 * FLOP_SAMPLE["stress"], 32 + 4 + 12          # minimal outputs: ray in, depth + rgb out
    return {"what": "cfg5 stress: 4096x4096 random rays against a 192x640 map, K=256, d_hidden=128, outputs rgb+depth; contiguous ray blocks",
            "rays_total": total_rays, "rays_per_gpu": n, "ms_per_step": ms, "rays_per_s": rps, "samples_per_s": rps * 256,
            "kernel": kern, "roofline": roofline(rps, flop_ray, bytes_ray, peaks)}


# ---------------------------------------------------------------------------------------------------------------------
# cfg3: KITTI-Raw training step (16 frames x 2048 patch rays, 4 views, 2 render views, shipped loss), ray shard
# ---------------------------------------------------------------------------------------------------------------------

def cfg3_train_step(dev, peaks, steps=10, with_reference=True):
    from behindthescenes_b200.loss import ReconstructionLoss
    from tests import grad_util as gu
    world = _world()
    n, Rr = 16, 2048                      # configs/exp_kitti_raw.yaml:11,59-61: batch 16, ray_batch_size 2048, 8x8 patches
    H, W = 370, 1226                      # KITTI-Raw frame (BASELINE config 3)
    wl = syn.cfg3_kitti_raw_train(H=H, W=W, n=n, lin_out_bias=-4.0, seed=61)
    # the rays of a training batch are 8x8 pixel patches (PatchRaySampler; 2048 rays = 32 patches per frame), drawn
    # independently per frame from the two loss views
    wl = syn.with_patch_rays(wl, H, W, 2, Rr // 64, seed=61)
    cpu_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in cpu_net.state_dict().items() if not k.startswith("encoder.")}
    net, wrapped = syn.build_model(wl, dev, state_dict=sd)
    net.encoder.feats.requires_grad_(True)
    imgs, Ks, poses, rays, u = (t.to(dev) for t in (wl.images, wl.Ks, wl.poses_c2w, wl.rays, wl.u))
    rgb_gt = torch.rand((n, Rr, 3), generator=torch.Generator().manual_seed(9)).to(dev)
    crit = ReconstructionLoss({"criterion": "l1+ssim", "invalid_policy": "weight_guided", "lambda_edge_aware_smoothness": 0.001})
    P = Rr // 64

    def ours():
        net.zero_grad(set_to_none=True)
        net.encode(imgs, Ks, poses, ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        out = wrapped(rays, u=u, want_weights=True, want_alphas=True, want_invalid=False, want_invalid_mass=True)["coarse"]
        coarse = {"rgb": out["rgb"].view(n, P, 8, 8, -1, 3), "depth": out["depth"].view(n, P, 8, 8),
                  "invalid_mass": out["invalid_mass"].view(n, P, 8, 8, -1)}
        loss = crit({"coarse": [coarse], "fine": [dict(coarse)], "rgb_gt": rgb_gt.view(n, P, 8, 8, 3)})[0]
        loss.backward()
        if world > 1:                     # one packed NCCL all-reduce of the field's gradients
            R.allreduce_gradients(net.mlp_coarse.parameters(), n * Rr)
        return loss

    ms_eager = timed_steps(ours, dev, steps, 3)
    for _ in range(2):
        ours()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(5):
        ours()
    host_ms = (time.perf_counter() - t0) * 1e3 / 5           # host time to enqueue a step (no sync inside)
    torch.cuda.synchronize(dev)
    # the same step captured in ONE CUDA graph (fold, render, loss, both backwards; the logged scalars stay on the device;
    # injected jitter u -- with in-kernel jitter the renderer's `seed_tensor` device word would be bumped inside the graph):
    # the ~7 ms of Python / launch overhead of the eager step disappear; the gradient all-reduce stays outside the graph
    crit.sync_scalars = False

    def local_step():
        net.zero_grad(set_to_none=True)
        net.encode(imgs, Ks, poses, ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        out = wrapped(rays, u=u, want_weights=True, want_alphas=True, want_invalid=False, want_invalid_mass=True)["coarse"]
        coarse = {"rgb": out["rgb"].view(n, P, 8, 8, -1, 3), "depth": out["depth"].view(n, P, 8, 8),
                  "invalid_mass": out["invalid_mass"].view(n, P, 8, 8, -1)}
        crit({"coarse": [coarse], "fine": [dict(coarse)], "rgb_gt": rgb_gt.view(n, P, 8, 8, 3)})[0].backward()

    replay, captured = graphed(local_step, dev)

    def graph_step():
        replay()
        if world > 1:
            R.allreduce_gradients(net.mlp_coarse.parameters(), n * Rr)

    ms_graph = timed_steps(graph_step, dev, steps, 2) if captured else None
    crit.sync_scalars = True
    ms = min(ms_eager, ms_graph) if ms_graph is not None else ms_eager
    res = {"what": f"cfg3 KITTI-Raw training step: {n} frames {H}x{W} x {Rr} rays per GPU in 8x8 patches (PatchRaySampler), K=64, 2 render views, fold + render + "
                   "shipped loss (l1+ssim, weight-guided mask, edge-aware smoothness) + backward (+ gradient all-reduce); encoder excluded",
           "rays_per_gpu": n * Rr, "ms_per_step": ms, "rays_per_s": world * n * Rr / ms * 1e3, "ms_per_step_eager": ms_eager,
           "ms_per_step_cuda_graph": ms_graph, "cuda_graph": captured, "host_enqueue_ms_eager": host_ms, "scaling": "weak"}
    # per-GPU algorithmic work of a step: the fold and its backward stream the feature maps (F read, G written; dG
    # zeroed + read, F read, dF written), the render forward / backward do the positional-code contraction per sample
    # (x3: the backward recomputes it and forms two gradient products) and move the per-ray tensors
    px = n * H * W
    # sparse fold: only the 64-pixel tiles the patches' epipolar lines cross are folded / differentiated; the dense part
    # that remains is the feature gradient (written once, zeros outside those tiles)
    hit = net._scene_cache.get(("sparse", 0))
    frac = float((hit[1]._sparse["mask"] != 0).float().mean().item()) if hit is not None else 1.0
    res["tiles_sampled_frac"] = frac
    flop = 3 * 2 * 64 * 64 * px * frac + 3 * 64 * (2 * 39 * 64 + 2 * 64) * n * Rr
    byts = (1 + 5 * frac) * px * 64 * 4 + n * Rr * (32 + 256 + 24 + 4 + 8 + 3 * 256 + 32 + 256 + 256 + 24 + 4 + 2 * 256)
    res["roofline"] = roofline(1e3 / ms, flop, byts, peaks)       # "rays" = steps here: totals per step
    res["roofline"]["note"] = ("totals per step per GPU: feature gradient written densely (px x C x 4 B) + 5 map-sized streams over the "
                               f"sampled tiles ({frac:.2f} of the map: F read twice, G written, d_G zeroed and read) + per-ray tensors")
    if with_reference and world == 1:
        from oracle import bts_oracle as orc
        from oracle import loss_oracle as lo
        from tests.golden_io import mlp_params
        sdg = {k: v.to(dev).requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.startswith("code_xyz")}
        feats_g = wl.feats.to(dev).requires_grad_(True)
        fcfg = orc.FieldConfig.from_model_conf(wl.model_conf)
        rcfg = orc.RenderConfig.from_renderer_conf(wl.renderer_conf)

        def ref():
            for t in list(sdg.values()) + [feats_g]:
                t.grad = None
            st = orc.encode_state(imgs, Ks, poses, feats_g, wl.ids_encoder, wl.ids_render)
            out = orc.render(st, mlp_params(sdg), fcfg, rcfg, rays, u=u, empty_feature=sdg.get("empty_feature"),
                             want_weights=True, want_alphas=True)["coarse"]
            inv = ((out["invalid"].to(torch.float32) * out["weights"].unsqueeze(-1)).sum(-2) > .9).all(-1).view(n, P, 8, 8)
            rgb6, gt5 = out["rgb"].view(n, P, 8, 8, -1, 3), rgb_gt.view(n, P, 8, 8, 3)
            loss = lo.patch_loss(rgb6, gt5, inv)[0] + lo.patch_loss(rgb6, gt5, inv)[0] + \
                0.001 * lo.smooth_loss(out["depth"].view(n, P, 8, 8), gt5, inv)[0]
            loss.backward()
            return loss

        try:
            ms_ref = timed_steps(ref, dev, 3, 1)
            l_ref, l_ours = float(ref().item()), float(ours().item())
            g_ours = net.mlp_coarse.lin_in.weight.grad.detach().clone()
            res["reference_eager_gpu"] = {"ms_per_step": ms_ref, "speedup": ms_ref / ms,
                                          "what": "the reference op chain + its loss with torch autograd (oracle port, eager) on this GPU"}
            res["parity"] = {"loss_ours": l_ours, "loss_reference": l_ref, "loss_rel_err": abs(l_ours - l_ref) / max(abs(l_ref), 1e-12),
                             "grad_rel_err_lin_in": gu.rel_err(g_ours.cpu(), sdg["mlp_coarse.lin_in.weight"].grad.cpu())}
        except Exception as e:                     # baseline only: never fail the bench because of it
            res["reference_eager_gpu"] = {"unavailable": repr(e)[:200]}
    return res


# ---------------------------------------------------------------------------------------------------------------------
# cfg4: RealEstate10K shape, 256x384, K = 128, d_hidden 32 + one ResnetFC block, distance code, 2 colour views
# ---------------------------------------------------------------------------------------------------------------------

def cfg4_re10k(dev, peaks, steps=10, with_parity=True):
    from behindthescenes_b200 import _lib
    world, rank = _world(), _rank()
    wl = syn.cfg4_re10k_nvs(H=256, W=384, K=128, lin_out_bias=-4.0, seed=14)
    cpu_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in cpu_net.state_dict().items() if not k.startswith("encoder.")}
    net, wrapped = syn.build_model(wl, dev, state_dict=sd)
    with torch.no_grad():
        net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    rays_all = wl.rays.to(dev)
    n_all = rays_all.shape[1]
    part = R.shard_rays(rays_all, rank, world, granularity=8).contiguous()

    def step():
        net.invalidate_scene()
        net._scene_struct()
        return wrapped(part, want_weights=True, want_alphas=True)

    with torch.no_grad():
        run, captured = graphed(step, dev)
        ms = timed_steps(run, dev, steps, 3)
    kern = _lib.load().bts_last_kernel().decode()
    rps = n_all / ms * 1e3
    bytes_ray = 32 + 4 * 32 + 12 * 2 + 4 + 12 * 2 + 4 * 128 * 2 + 2 * 4 * 128
    res = {"what": "cfg4 RealEstate10K shape: novel view 256x384 (98304 rays) split over the ranks, K=128, C=32, d_hidden=32 + 1 ResnetFC block, "
                   "distance code, 2 colour views, outputs rgb+depth+invalid+weights+alphas; step = fold + pack + render",
           "rays_total": n_all, "rays_per_gpu": part.shape[1], "ms_per_frame": ms, "rays_per_s": rps, "kernel": kern, "cuda_graph": captured,
           "roofline": roofline(rps, 128 * FLOP_SAMPLE["re10k"], bytes_ray, peaks)}
    if with_parity and world == 1:
        from tests import parity_util as pu
        n_s = 8192
        with torch.no_grad():
            got = wrapped(rays_all[:, :n_s].contiguous(), u=wl.u[:n_s].to(dev), want_weights=True, want_alphas=True)["coarse"]
        sub = syn.Workload(**{**wl.__dict__, "rays": wl.rays[:, :n_s].contiguous(), "u": wl.u[:n_s].contiguous()})
        ref = pu.oracle_render(sub, sd, want_weights=True, want_alphas=True)
        d, dr = got["depth"].cpu(), ref["depth"]
        same = (got["invalid"].cpu() == ref["invalid"]).flatten(2).all(-1)
        res["parity"] = {"rays": n_s, "depth_abs_rel": ((d - dr).abs() / dr)[same].mean().item(),
                         "max_abs_weights": (got["weights"].cpu() - ref["weights"])[same].abs().max().item(),
                         "flag_flip_rays": int((~same).sum().item())}
    return res


# ---------------------------------------------------------------------------------------------------------------------
# What a user's frame costs with an encoder in front: encoder stand-in + encode() glue + render through the public API
# ---------------------------------------------------------------------------------------------------------------------

class _EncoderStandIn(torch.nn.Module):
    """A stand-in with the SHAPE of the reference's backbone (``models/common/backbones/monodepth2.py``: ResNet-50 trunk,
    U-Net decoder with skip connections up to full resolution, d_out = 64 feature channels), random weights, written here
    from torchvision's ResNet -- the encoder itself is out of scope and the reference is not on the GPU box;
    this only prices ``encode()`` so that the render time can be put next to it."""

    def __init__(self, d_out=64):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet50(weights=None)
        self.stem = torch.nn.Sequential(r.conv1, r.bn1, r.relu)
        self.pool, self.layers = r.maxpool, torch.nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])
        enc = [64, 256, 512, 1024, 2048]
        dec = [max(d_out, c) for c in (16, 32, 64, 128, 256)]
        conv = lambda i, o: torch.nn.Sequential(torch.nn.ReflectionPad2d(1), torch.nn.Conv2d(i, o, 3), torch.nn.ELU(inplace=True))  # noqa: E731
        self.up0 = torch.nn.ModuleList([conv(enc[4] if i == 4 else dec[i + 1], dec[i]) for i in range(5)])
        self.up1 = torch.nn.ModuleList([conv(dec[i] + (enc[i - 1] if i > 0 else 0), dec[i]) for i in range(5)])
        self.out = torch.nn.Conv2d(dec[0], d_out, 1)
        self.latent_size, self.scales = d_out, [0]

    def forward(self, x):
        f = [self.stem(x)]
        y = self.pool(f[0])
        for layer in self.layers:
            y = layer(y)
            f.append(y)
        y = f[4]
        for i in range(4, -1, -1):
            y = torch.nn.functional.interpolate(self.up0[i](y), scale_factor=2, mode="nearest")
            if i > 0:
                y = torch.cat((y, f[i - 1]), 1)
            y = self.up1[i](y)
        return [self.out(y)]


def frame_with_encoder(wl, sd, dev, steps=10):
    """encode() (stand-in encoder + the path's glue) + render of the cfg2 stereo frame through the drop-in classes."""
    net, wrapped = syn.build_model(wl, dev, state_dict=sd)
    net.encoder = _EncoderStandIn(net.encoder.latent_size).to(dev).eval()
    imgs, Ks, poses, rays = (t.to(dev) for t in (wl.images, wl.Ks, wl.poses_c2w, wl.rays))
    wants = dict(want_weights=True, want_alphas=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_enc = t_all = 0.0
    with torch.no_grad():
        for i in range(-3, steps):
            ev[0].record()
            net.encode(imgs, Ks, poses, ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
            ev[1].record()
            wrapped(rays, **wants)
            ev[2].record()
            torch.cuda.synchronize(dev)
            if i >= 0:
                t_enc += ev[0].elapsed_time(ev[1])
                t_all += ev[0].elapsed_time(ev[2])
    n_rays = rays.shape[1]
    ms_enc, ms_all = _max_over_ranks(t_enc / steps, dev), _max_over_ranks(t_all / steps, dev)
    return {"what": "cfg2 stereo frame with an encoder in front: net.encode() with a ResNet-50 + U-Net stand-in of the reference backbone's shape "
                    "(random weights, fp32 eager cuDNN; NOT the reference's encoder, which is out of scope and not on this box) + fold + pack + render",
            "ms_encode_standin": ms_enc, "ms_fold_pack_render": ms_all - ms_enc, "ms_per_frame": ms_all,
            "rays_per_s": n_rays * _world() / (ms_all * 1e-3), "render_share": (ms_all - ms_enc) / ms_all}
