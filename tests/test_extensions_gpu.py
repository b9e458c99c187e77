"""GPU: the §8f "next" rows -- fused consumers of the per-sample outputs, in-kernel ray generation, z-depth."""
import pytest
import torch

from behindthescenes_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def _model(wl):
    net, wrapped = syn.build_model(wl, "cuda")
    with torch.no_grad():
        net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    return net, wrapped


@pytest.mark.parametrize("precision", [0, 2])
def test_invalid_mass_equals_the_reduction_of_the_per_sample_outputs(precision):
    """loss.py:107 / inference_setup.py:192 compute (invalid * weights).sum(-2); the kernel can emit it directly."""
    wl = syn.cfg3_kitti_raw_train(H=48, W=160, n=2, lin_out_bias=-4.0, seed=51, ray_subsample=4096)
    net, wrapped = _model(wl)
    wrapped.renderer.precision = precision
    with torch.no_grad():
        torch.manual_seed(1)
        a = wrapped(wl.rays.cuda(), want_weights=True, want_invalid_mass=True)["coarse"]
        torch.manual_seed(1)
        b = wrapped(wl.rays.cuda(), want_invalid=False, want_invalid_mass=True)["coarse"]
    want = (a["invalid"] * a["weights"].unsqueeze(-1)).sum(-2)
    assert a["invalid_mass"].shape == want.shape == (2, 4096, 2)
    assert (a["invalid_mass"] - want).abs().max().item() <= 2e-6
    assert "invalid" not in b and torch.equal(b["invalid_mass"], a["invalid_mass"]) and torch.equal(b["rgb"], a["rgb"])
    assert float(want.max()) > 0.5          # the fixture does contain rays that leave the frustum


@pytest.mark.parametrize("precision", [0, 2])
def test_in_kernel_ray_generation_matches_ray_tensor_path(precision):
    """render_views(poses, Ks, H, W) == forward(gen_rays(...)) (models/common/util/util.py:244-273), plus z-depth
    (utils/projection_operations.py:4-16)."""
    H, W = 24, 80
    wl = syn.cfg3_kitti_raw_train(H=H, W=W, n=2, lin_out_bias=-4.0, seed=52)
    net, wrapped = _model(wl)
    wrapped.renderer.precision = precision
    views = [0, 3]
    poses = wl.poses_c2w[:, views]                                   # (n, 2, 4, 4)
    Ks = wl.Ks[:, views]
    f = torch.stack((Ks[0, :, 0, 0], Ks[0, :, 1, 1]), -1)
    c = torch.stack((Ks[0, :, 0, 2], Ks[0, :, 1, 2]), -1)
    rays = torch.stack([syn.gen_rays(poses[b], W, H, 3.0, 80.0, f, c).reshape(-1, 8) for b in range(2)], 0)
    K = wl.renderer_conf["n_coarse"]
    u = torch.rand((rays.shape[0] * rays.shape[1], K), generator=torch.Generator().manual_seed(3)).cuda()
    wants = dict(want_weights=True, want_alphas=True, want_z_samps=True)
    with torch.no_grad():
        a = wrapped(rays.cuda(), u=u, **wants)["coarse"]
        b = wrapped.render_views(poses.cuda(), Ks.cuda(), H, W, 3.0, 80.0, u=u, want_depth_z=True, **wants)["coarse"]
    assert b["depth"].shape == (2, 2 * H * W) and b["depth_z"].shape == (2, 2 * H * W)
    assert torch.equal(a["z_samps"], b["z_samps"])
    same = (a["invalid"] == b["invalid"]).flatten(2).all(-1)
    assert same.float().mean().item() > 0.98                         # knife-edge flags may flip (ulp-level ray differences)
    rel = ((a["depth"] - b["depth"]).abs() / a["depth"])[same]
    assert rel.mean().item() <= 1e-5 and rel.max().item() <= 1e-4
    assert (a["weights"] - b["weights"])[same].abs().max().item() <= 1e-4
    assert (a["rgb"] - b["rgb"])[same].abs().max().item() <= 1e-4
    # z-depth: depth * cam_z / |cam|, cam = K^-1 (x, y, 1)
    xs = torch.linspace(-1, 1, W).view(1, 1, 1, W).expand(2, 2, H, W)
    ys = torch.linspace(-1, 1, H).view(1, 1, H, 1).expand(2, 2, H, W)
    pts = torch.stack((xs, ys, torch.ones_like(xs)), 2).view(2, 2, 3, -1)
    cam = (torch.inverse(Ks) @ pts).view(2, 2, 3, H, W)
    fac = (cam[:, :, 2] / cam.norm(dim=2)).reshape(2, -1)
    assert torch.allclose(b["depth_z"].cpu(), b["depth"].cpu() * fac, rtol=2e-6, atol=1e-6)


def test_output_diet_trains_but_in_kernel_rays_are_inference_only():
    """want_invalid=False / want_invalid_mass=True are part of the autograd node (non-differentiable outputs, as in the
    reference where loss.py:105-108 only thresholds them); in-kernel ray generation / depth_z stay inference-only."""
    wl = syn.cfg2_kitti360_eval(H=16, W=48, ray_subsample=64)
    net, wrapped = _model(wl)
    out = wrapped(wl.rays.cuda(), u=wl.u.cuda(), want_weights=True, want_invalid=False, want_invalid_mass=True)["coarse"]
    assert "invalid" not in out and out["rgb"].requires_grad and not out["invalid_mass"].requires_grad
    with torch.no_grad():
        ref = wrapped(wl.rays.cuda(), u=wl.u.cuda(), want_weights=True)["coarse"]
    mass = (ref["invalid"] * ref["weights"].unsqueeze(-1)).sum(-2)
    assert torch.allclose(out["invalid_mass"], mass, atol=1e-5) and torch.equal(out["rgb"].detach(), ref["rgb"])
    out["rgb"].sum().backward()
    assert net.mlp_coarse.lin_in.weight.grad is not None
    with pytest.raises(NotImplementedError):
        wrapped.render_views(wl.poses_c2w.cuda()[:, :1], wl.Ks.cuda()[:, :1], 16, 48, 3.0, 80.0)   # grad enabled + trainable parameters
