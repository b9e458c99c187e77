"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (``/root/reference``) on small synthetic
frustums.  TEST INFRASTRUCTURE ONLY; runs in the build container (the reference does not travel to the GPU box).

    python -m oracle.make_goldens            # rewrites every fixture

Each fixture stores all inputs (images, intrinsics, poses, encoder features, MLP weights, rays, jitter ``u``), the
configuration (JSON) and the reference outputs of ``_RenderWrapper.forward`` (models/common/render/nerf.py:19-45) with
every optional output requested, plus ``BTSNet.forward`` point-query outputs (models/bts/model/models_bts.py:266-338).
The jitter is injected by seeding torch right before the reference's ``torch.rand_like`` (nerf.py:116): on CPU
``rand_like`` of a (B,K) tensor consumes the same stream as ``torch.rand((B,K))`` (checked below, bit-exact z).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from behindthescenes_b200 import synthetic as syn  # noqa: E402
from oracle import ref_import  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def fixtures():
    P = syn._pose
    out = []
    out.append(syn.cfg2_kitti360_eval(H=16, W=48, K=64, stereo=True, learn_empty=True, lin_out_bias=0.0, seed=1,
                                      ray_subsample=192))
    out[-1].name = "cfg2_small_bias0"
    out.append(syn.cfg2_kitti360_eval(H=16, W=48, K=64, stereo=True, learn_empty=True, lin_out_bias=-4.0, seed=2,
                                      ray_subsample=192))
    out[-1].name = "cfg2_small_bias4"
    out.append(syn.cfg2_kitti360_eval(H=16, W=48, K=32, stereo=False, learn_empty=True, lin_out_bias=-4.0, seed=3,
                                      ray_subsample=192))
    out[-1].name = "cfg1_small_K32"
    out.append(syn.cfg3_kitti_raw_train(H=20, W=64, K=64, n=2, seed=4, lin_out_bias=-4.0, ray_subsample=96))
    out[-1].name = "cfg3_small_n2"
    out.append(syn.cfg4_re10k_nvs(H=24, W=36, K=128, seed=5, lin_out_bias=-4.0, ray_subsample=96))
    out[-1].name = "cfg4_small"
    # edge cases: linear depth sampling, linear depth code, white background, no alpha cap, rays that leave the
    # frustum and pass behind the encoder camera (clamped z -> huge |uv|), K=32, learn_empty on
    mc = syn.kitti_model_conf(learn_empty=True, z_near=0.5, z_far=20.0, inv_z=False)
    rc = syn.kitti_renderer_conf(32, lindisp=False, hard_alpha_cap=False, white_bkgd=True)
    out.append(syn.make_workload("edge_linear_white", H=16, W=48, poses=[P(), P(tx=0.3, yaw_deg=8.0)], model_conf=mc,
                                 renderer_conf=rc, ids_encoder=[0], ids_render=[0, 1],
                                 ray_poses=[P(tx=1.0, tz=4.0, yaw_deg=160.0), P(ty=-0.5, tz=-2.0, yaw_deg=-35.0, pitch_deg=10.0)],
                                 ray_subsample=192, seed=6, lin_out_bias=-2.0))
    # distance code with inverse depth and learn_empty, 1 block, Hd=64
    mc = syn.kitti_model_conf(learn_empty=True, n_blocks=1, code_mode="distance")
    rc = syn.kitti_renderer_conf(64)
    out.append(syn.make_workload("dist_block1", H=16, W=48, poses=[P(), P(tx=0.54)], model_conf=mc, renderer_conf=rc,
                                 ids_encoder=[0], ids_render=[1], ray_poses=[P(tx=0.2, tz=-0.5, yaw_deg=12.0)],
                                 ray_subsample=128, seed=7, lin_out_bias=-3.0))
    return out


GRAD_FIXTURES = ("cfg2_small_bias4", "cfg3_small_n2", "cfg4_small")


def probe_loss(coarse):
    """Deterministic scalar of the differentiable outputs (used for the gradient fixtures and their tests):
    sum_t <t, sin(0.37 i + phase_t)> over rgb, depth (x0.1), weights, alphas."""
    total = 0.0
    for name, phase, scale in (("rgb", 0.1, 1.0), ("depth", 0.2, 0.1), ("weights", 0.3, 1.0), ("alphas", 0.4, 1.0)):
        t = coarse[name]
        coef = torch.sin(0.37 * torch.arange(t.numel(), dtype=torch.float32, device=t.device) + phase).view_as(t)
        total = total + scale * (t * coef).sum()
    return total


def reference_gradients(ref, wl: syn.Workload, sd):
    """Gradients of probe_loss through the UNMODIFIED reference (torch autograd, like utils/base_trainer.py:297)."""
    net = ref_import.make_reference_net(ref, wl.model_conf, wl.feats)
    net.load_state_dict(sd, strict=False)
    renderer = ref.NeRFRenderer.from_conf(wl.renderer_conf)
    wrapped = renderer.bind_parallel(net, gpus=None).eval()
    net.encode(wl.images, wl.Ks, wl.poses_c2w, ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    net.set_scale(0)
    torch.manual_seed(12345)
    out = wrapped(wl.rays, want_weights=True, want_alphas=True)["coarse"]
    probe_loss(out).backward()
    g = {"mlp_coarse.lin_in.weight": net.mlp_coarse.lin_in.weight.grad, "mlp_coarse.lin_in.bias": net.mlp_coarse.lin_in.bias.grad,
         "mlp_coarse.lin_out.weight": net.mlp_coarse.lin_out.weight.grad, "mlp_coarse.lin_out.bias": net.mlp_coarse.lin_out.bias.grad,
         "feats": net.encoder.feats.grad}
    if getattr(net, "learn_empty", False):
        g["empty_feature"] = net.empty_feature.grad
    for bi, blk in enumerate(net.mlp_coarse.blocks):
        for nm, prm in (("fc_0.weight", blk.fc_0.weight), ("fc_0.bias", blk.fc_0.bias), ("fc_1.weight", blk.fc_1.weight),
                        ("fc_1.bias", blk.fc_1.bias)):
            g[f"mlp_coarse.blocks.{bi}.{nm}"] = prm.grad
    return {k: v.detach().clone() for k, v in g.items()}


def run_reference(ref, wl: syn.Workload):
    torch.manual_seed(wl.weight_seed)
    net = ref_import.make_reference_net(ref, wl.model_conf, wl.feats)
    g = torch.Generator().manual_seed(wl.weight_seed + 100)
    with torch.no_grad():
        net.mlp_coarse.lin_out.bias.fill_(wl.lin_out_bias)
        for blk in net.mlp_coarse.blocks:   # reference zero-inits fc_1 (resnetfc.py:40): randomise to exercise it
            blk.fc_1.weight.copy_(torch.randn(blk.fc_1.weight.shape, generator=g) * (2.0 / blk.fc_1.weight.shape[1]) ** 0.5 * 0.5)
            blk.fc_0.bias.copy_(torch.randn(blk.fc_0.bias.shape, generator=g) * 0.1)
            blk.fc_1.bias.copy_(torch.randn(blk.fc_1.bias.shape, generator=g) * 0.1)
        net.mlp_coarse.lin_in.bias.copy_(torch.randn(net.mlp_coarse.lin_in.bias.shape, generator=g) * 0.1)
    renderer = ref.NeRFRenderer.from_conf(wl.renderer_conf)
    wrapped = renderer.bind_parallel(net, gpus=None).eval()
    with torch.no_grad():
        net.encode(wl.images, wl.Ks, wl.poses_c2w, ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        net.set_scale(0)
        # inject u: reseed so that rand_like (nerf.py:116) reproduces wl.u
        B, K = wl.u.shape
        torch.manual_seed(12345)
        u_check = torch.rand((B, K))
        torch.manual_seed(12345)
        out = wrapped(wl.rays, want_weights=True, want_alphas=True, want_z_samps=True, want_rgb_samps=True)
        coarse = out["coarse"]
        # point query on the first ray points + a few hand-placed points (behind camera, on the optical axis)
        z = coarse["z_samps"]
        pts = (wl.rays[:, :, None, :3] + z.unsqueeze(-1) * wl.rays[:, :, None, 3:6])[:, ::7].reshape(wl.rays.shape[0], -1, 3)
        extra = torch.tensor([[0.0, 0.0, -1.0], [0.0, 0.0, 5.0], [100.0, 0.0, 1.0], [0.0, 0.0, 1e-4]]).view(1, -1, 3)
        pts = torch.cat((pts[:, :500], extra.expand(pts.shape[0], -1, -1)), 1).contiguous()
        q_rgb, q_inv, q_sig = net(pts)
        d_rgb, d_inv, d_sig = net(pts, only_density=True)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items() if not k.startswith("encoder.")}
    return u_check, coarse, sd, dict(pts=pts, q_rgb=q_rgb, q_invalid=q_inv, q_sigma=q_sig, d_rgb=d_rgb,
                                     d_invalid=d_inv, d_sigma=d_sig)


def main():
    ref = ref_import.load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for wl in fixtures():
        u, coarse, sd, q = run_reference(ref, wl)
        # the stratified depths the reference produced must be reproducible from u (bit-exact)
        from oracle import bts_oracle as orc
        z_from_u = orc.stratified_depths(wl.rays.reshape(-1, 8), wl.renderer_conf["n_coarse"], wl.renderer_conf["lindisp"], u)
        assert torch.equal(z_from_u.view_as(coarse["z_samps"]), coarse["z_samps"]), "jitter injection is not bit-exact"
        arrays = {
            "images": wl.images, "Ks": wl.Ks, "poses_c2w": wl.poses_c2w, "feats": wl.feats, "rays": wl.rays, "u": u,
        }
        for k, v in sd.items():
            arrays["sd/" + k] = v
        for k, v in coarse.items():
            arrays["out/" + k] = v
        for k, v in q.items():
            arrays["query/" + k] = v
        if wl.name in GRAD_FIXTURES:
            for k, v in reference_gradients(ref, wl, sd).items():
                arrays["grad/" + k] = v
        meta = {"name": wl.name, "model_conf": wl.model_conf, "renderer_conf": wl.renderer_conf,
                "ids_encoder": wl.ids_encoder, "ids_render": wl.ids_render,
                "reference": "Brummi/BehindTheScenes @ /root/reference (unmodified; stubs: dotmap, lpips, omegaconf; "
                             "encoder replaced by a feature stand-in)", "torch": torch.__version__}
        path = os.path.join(GOLDEN_DIR, wl.name + ".npz")
        np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                            **{k: v.detach().cpu().numpy() for k, v in arrays.items()})
        inv = coarse["invalid"]
        print(f"{wl.name}: rays {tuple(wl.rays.shape)} K={wl.renderer_conf['n_coarse']} depth "
              f"[{coarse['depth'].min():.2f},{coarse['depth'].max():.2f}] invalid frac {inv.mean():.3f} "
              f"-> {os.path.getsize(path) / 1e3:.0f} kB")


def make_loss_golden(ref):
    """Patch photometric loss fixture from the reference's own functions (models/bts/model/loss.py)."""
    import importlib
    loss_mod = importlib.import_module("models.bts.model.loss")
    g = torch.Generator().manual_seed(77)
    n, pc, h, w, nv, c = 2, 6, 8, 8, 2, 3
    gt = torch.rand((n, pc, h, w, c), generator=g)
    pred = (gt.unsqueeze(-2) + 0.15 * torch.randn((n, pc, h, w, nv, c), generator=g)).clamp(0, 1)
    pred[0, 0] = gt[0, 0].unsqueeze(-2)                       # a perfectly reconstructed patch (SSIM clamp active)
    pred[0, 1, :, :, 0] = 0.0                                 # a black view
    invalid = torch.rand((n, pc, h, w), generator=g) > 0.8
    pred = pred.clone().requires_grad_(True)
    errors = loss_mod.compute_errors_l1ssim(pred, gt.unsqueeze(-2))            # (n,pc,h,w,nv,1)   loss.py:154
    per_ray = errors.amin(-2) * (1 - invalid.unsqueeze(-1).to(torch.float32))   # :155, :160-161
    loss = per_ray.mean()                                                       # :166
    loss.backward()
    os.makedirs(os.path.join(GOLDEN_DIR, "loss"), exist_ok=True)
    path = os.path.join(GOLDEN_DIR, "loss", "l1ssim.npz")
    np.savez_compressed(path, pred=pred.detach().numpy(), gt=gt.numpy(), invalid=invalid.numpy(),
                        errors=errors.detach().numpy(), per_ray=per_ray.detach().numpy(), loss=loss.detach().numpy(),
                        grad_pred=pred.grad.numpy())
    print(f"loss_l1ssim: loss {loss.item():.6f} -> {os.path.getsize(path) / 1e3:.0f} kB")


RECON_CONFIGS = {
    "shipped": {"criterion": "l1+ssim", "invalid_policy": "weight_guided", "lambda_edge_aware_smoothness": 0.001},   # configs/exp_*.yaml
    "all_terms": {"criterion": "l1+ssim", "invalid_policy": "strict", "lambda_fine": 0.5, "lambda_entropy": 0.1,
                  "lambda_depth_reg": 0.01, "lambda_alpha_reg": 0.05, "lambda_surfaceness_reg": 0.02,
                  "lambda_edge_aware_smoothness": 0.01, "lambda_depth_smoothness": 0.03, "alpha_reg_fraction": 0.25},
    "unmasked_slice": {"criterion": "l1+ssim", "invalid_policy": "none", "lambda_alpha_reg": 0.05, "alpha_reg_reduction": "slice",
                       "alpha_reg_fraction": 0.25, "lambda_edge_aware_smoothness": 0.002},
}


def recon_loss_inputs():
    """Seeded stand-in for the trainer's ``data`` dict after PatchRaySampler.reconstruct (models/bts/trainer.py:236-258)."""
    g = torch.Generator().manual_seed(123)
    n, pc, h, w, nv, c, K = 2, 4, 8, 8, 2, 3, 8
    gt = torch.rand((n, pc, h, w, c), generator=g)
    rgb = (gt.unsqueeze(-2) + 0.1 * torch.randn((n, pc, h, w, nv, c), generator=g)).clamp(0, 1)
    depth = 3 + 60 * torch.rand((n, pc, h, w), generator=g)
    depth[0, 0, :2] = 90.0                                    # beyond the clamp of edge_aware_smoothness
    depth[0, 1] = 20.0                                        # a flat patch: |d - d'| = 0 everywhere
    alphas = torch.rand((n, pc, h, w, K), generator=g) * 0.5
    weights = torch.softmax(torch.randn((n, pc, h, w, K), generator=g) * 2, -1)
    invalid = torch.rand((n, pc, h, w, K, nv), generator=g) > 0.6
    invalid[:, :, :3] = True                                  # rays whose whole mass is invalid in every view
    invalid[:, :, 3:5, :, :, 1] = False
    return dict(rgb=rgb, depth=depth, alphas=alphas, weights=weights, invalid=invalid.to(torch.float32), rgb_gt=gt)


def make_recon_loss_golden(ref):
    """ReconstructionLoss.__call__ of the unmodified reference (models/bts/model/loss.py:43-293) on the seeded data dict."""
    import importlib
    loss_mod = importlib.import_module("models.bts.model.loss")
    out = {}
    base = recon_loss_inputs()
    for k, v in base.items():
        out["in/" + k] = v.numpy()
    for name, conf in RECON_CONFIGS.items():
        t = {k: v.clone() for k, v in base.items()}
        for k in ("rgb", "depth", "alphas"):
            t[k].requires_grad_(True)
        coarse = {k: t[k] for k in ("rgb", "depth", "alphas", "weights", "invalid")}
        data = {"coarse": [coarse], "fine": [dict(coarse)], "rgb_gt": t["rgb_gt"]}      # trainer.py:246-247
        loss, loss_dict = loss_mod.ReconstructionLoss(conf)(data)
        loss.backward()
        out[f"{name}/loss"] = loss.detach().numpy()
        out[f"{name}/dict"] = np.frombuffer(json.dumps({k: float(v) for k, v in loss_dict.items()}).encode(), dtype=np.uint8)
        for k in ("rgb", "depth", "alphas"):
            out[f"{name}/grad_{k}"] = (t[k].grad if t[k].grad is not None else torch.zeros_like(t[k])).numpy()
        print(f"recon_loss[{name}]: loss {loss.item():.6f} {loss_dict}")
    os.makedirs(os.path.join(GOLDEN_DIR, "loss"), exist_ok=True)
    path = os.path.join(GOLDEN_DIR, "loss", "reconstruction_loss.npz")
    np.savez_compressed(path, **out)
    print(f"-> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
    make_loss_golden(ref_import.load_reference())
    make_recon_loss_golden(ref_import.load_reference())
