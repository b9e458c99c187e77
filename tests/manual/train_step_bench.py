"""Training-shaped step (BASELINE config 3 flavour): forward + backward of the render path on patch rays.
Ours (fused fwd + bts_render_bwd) vs the reference op chain with torch autograd on the same GPU.  Run under gpurun."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from behindthescenes_b200 import synthetic as syn
from tests import grad_util as gu

n, R = 16, 2048                       # configs/exp_kitti_raw.yaml: batch_size 16, ray_batch_size 2048, patch 8x8
H, W = 192, 640
wl = syn.cfg3_kitti_raw_train(H=H, W=W, n=n, lin_out_bias=-4.0, seed=61, ray_subsample=R)
WORLD, RANK = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
if WORLD > 1:      # ray shard, weak scaling: every rank renders its own 16 x 2048 patch rays of the replicated frames
    import torch.distributed as dist
    from behindthescenes_b200 import rays as R_
    dist.init_process_group("nccl", device_id=dev)
cpu_net, _ = syn.build_model(wl, "cpu")
sd = {k: v.clone() for k, v in cpu_net.state_dict().items() if not k.startswith("encoder.")}
net, wrapped = syn.build_model(wl, dev, state_dict=sd)
net.encoder.feats.requires_grad_(True)
imgs, Ks, poses, rays, u = (t.to(dev) for t in (wl.images, wl.Ks, wl.poses_c2w, wl.rays, wl.u))
wants = dict(want_weights=True, want_alphas=True, want_rgb_samps=True)   # models/bts/trainer.py:245
LOSS = os.environ.get("LOSS", "probe")      # "l1ssim": the reference's photometric patch loss on both arms
rgb_gt = torch.rand((n, R, 3), generator=torch.Generator().manual_seed(9)).to(dev)


def ours():
    net.zero_grad(set_to_none=True)
    net.encode(imgs, Ks, poses, ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    if LOSS == "shipped":       # the loss of every configs/exp_*.yaml through the drop-in ReconstructionLoss
        from behindthescenes_b200.loss import ReconstructionLoss
        out = wrapped(rays, u=u, want_weights=True, want_alphas=True, want_invalid=False, want_invalid_mass=True)["coarse"]
        coarse = {"rgb": out["rgb"].view(n, R // 64, 8, 8, -1, 3), "depth": out["depth"].view(n, R // 64, 8, 8),
                  "invalid_mass": out["invalid_mass"].view(n, R // 64, 8, 8, -1)}
        crit = ReconstructionLoss({"criterion": "l1+ssim", "invalid_policy": "weight_guided", "lambda_edge_aware_smoothness": 0.001})
        crit({"coarse": [coarse], "fine": [dict(coarse)], "rgb_gt": rgb_gt.view(n, R // 64, 8, 8, 3)})[0].backward()
        if WORLD > 1:      # one packed NCCL all-reduce of the field's gradients (the encoder's would ride in the same bucket)
            R_.allreduce_gradients(net.mlp_coarse.parameters(), n * R)
        return
    if LOSS == "l1ssim":
        from behindthescenes_b200.loss import patch_photometric_loss
        out = wrapped(rays, u=u, want_weights=True, want_alphas=True, want_invalid=False, want_invalid_mass=True)["coarse"]
        patch_photometric_loss(out["rgb"], rgb_gt, invalid_mass=out["invalid_mass"], patch=(8, 8))[0].backward()
        return
    out = wrapped(rays, u=u, **wants)["coarse"]
    gu.probe_loss(out).backward()


HOST_MS = {}


def timeit(fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    HOST_MS[fn.__name__] = (time.perf_counter() - t0) * 1e3 / steps      # host time to ENQUEUE a step (no sync inside)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


ms_ours = timeit(ours, steps=int(os.environ.get("STEPS", 10)), warm=int(os.environ.get("WARM", 3)))
if WORLD > 1:
    t = torch.tensor([ms_ours], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if RANK == 0:
        print(json.dumps({"workload": f"cfg3-shaped training step, ray shard x{WORLD} (weak: {n}x{R} rays per rank), loss={LOSS}, "
                                      "gradient all-reduce over NCCL every step", "n_gpus": WORLD, "ms_per_step_max_over_ranks": t.item(),
                          "rays_per_s": WORLD * n * R / t.item() * 1e3}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0)
if os.environ.get("PROFILE_HOST"):       # where the host time of a step goes
    import cProfile, pstats, io
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        ours()
    pr.disable()
    torch.cuda.synchronize()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(45)
    print(buf.getvalue())
    sys.exit(0)
if os.environ.get("ONLY_OURS"):          # launch-list runs under ncu
    print(json.dumps({"ms_ours": ms_ours, "host_enqueue_ms": HOST_MS.get("ours"), "loss": LOSS}))
    sys.exit(0)
# reference op chain with autograd on the GPU
from oracle import bts_oracle as orc
from tests.golden_io import mlp_params
sdg = {k: v.to(dev).requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.startswith("code_xyz")}
feats_g = wl.feats.to(dev).requires_grad_(True)
fcfg = orc.FieldConfig.from_model_conf(wl.model_conf)
rcfg = orc.RenderConfig.from_renderer_conf(wl.renderer_conf)


def ref():
    for t in list(sdg.values()) + [feats_g]:
        t.grad = None
    st = orc.encode_state(imgs, Ks, poses, feats_g, wl.ids_encoder, wl.ids_render)
    out = orc.render(st, mlp_params(sdg), fcfg, rcfg, rays, u=u, empty_feature=sdg.get("empty_feature"), **wants)["coarse"]
    if LOSS == "shipped":       # what ReconstructionLoss.__call__ runs for the shipped configs (fine = copy of coarse)
        from oracle import loss_oracle as lo
        inv = ((out["invalid"].to(torch.float32) * out["weights"].unsqueeze(-1)).sum(-2) > .9).all(-1).view(n, R // 64, 8, 8)
        rgb6, gt5 = out["rgb"].view(n, R // 64, 8, 8, -1, 3), rgb_gt.view(n, R // 64, 8, 8, 3)
        loss = lo.patch_loss(rgb6, gt5, inv)[0] + lo.patch_loss(rgb6, gt5, inv)[0] + \
            0.001 * lo.smooth_loss(out["depth"].view(n, R // 64, 8, 8), gt5, inv)[0]
        loss.backward()
        return
    if LOSS == "l1ssim":
        from oracle import loss_oracle as lo
        inv = ((out["invalid"].to(torch.float32) * out["weights"].unsqueeze(-1)).sum(-2) > .9).all(-1)      # loss.py:107
        lo.patch_loss(out["rgb"].view(n, R // 64, 8, 8, -1, 3), rgb_gt.view(n, R // 64, 8, 8, 3), inv.view(n, R // 64, 8, 8))[0].backward()
        return
    gu.probe_loss(out).backward()


ms_ref = timeit(ref, steps=5, warm=2)
g_ours = net.mlp_coarse.lin_in.weight.grad.detach().clone()
g_ref = sdg["mlp_coarse.lin_in.weight"].grad
print(json.dumps({"workload": f"cfg3-shaped training step: n={n} frames {H}x{W}, {R} rays each, K=64, Vc=2, fwd+bwd of the render path (encoder excluded), loss=" + LOSS,
                  "ms_ours": ms_ours, "host_enqueue_ms_ours": HOST_MS.get("ours"), "ms_reference_eager_gpu": ms_ref, "host_enqueue_ms_reference": HOST_MS.get("ref"), "speedup": ms_ref / ms_ours,
                  "rays_per_s_ours": n * R / ms_ours * 1e3, "grad_rel_err_lin_in": gu.rel_err(g_ours.cpu(), g_ref.cpu())}))
