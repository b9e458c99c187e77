"""GPU: bts_render_bwd (through torch autograd on the drop-in classes) against the gradients of the UNMODIFIED
reference (fixtures) and against autograd through the oracle at a training-shaped size."""
import pytest
import torch

from behindthescenes_b200 import synthetic as syn
from tests import grad_util as gu
from tests import parity_util as pu
from tests.golden_io import golden_names, load_golden

pytestmark = pytest.mark.gpu
# Relative L2 error per gradient tensor.  Well-conditioned workloads measure 3e-7 .. 4e-6 on B200 (r02) and are held to
# 5e-5 -- far below what a sign or term error in even a small contribution would produce.  Where the fp32 reference gradient
# itself is ill conditioned (ReLU-mask / frustum-flag flips of single samples under ulp-level forward differences; linear
# depth sampling from z_near = 0.5), the bound is 4x the measured distance between the fp32 and the float64 evaluation of
# the reference (tests/grad_util.py::grad_tolerances) -- two correct fp32 evaluations cannot agree better than that.
GRAD_REL_TOL = 5e-5


def cuda_gradients(wl, sd, precision=0, use_u=True, loss_fn=gu.probe_loss):
    net, wrapped = syn.build_model(wl, "cuda", state_dict=sd)
    wrapped.renderer.precision = precision
    net.encoder.feats.requires_grad_(True)
    net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    out = wrapped(wl.rays.cuda(), want_weights=True, want_alphas=True, u=wl.u.cuda() if use_u else None)["coarse"]
    assert out["rgb"].requires_grad and out["depth"].requires_grad and not out["invalid"].requires_grad
    loss_fn(out).backward()
    g = {"mlp_coarse.lin_in.weight": net.mlp_coarse.lin_in.weight.grad, "mlp_coarse.lin_in.bias": net.mlp_coarse.lin_in.bias.grad,
         "mlp_coarse.lin_out.weight": net.mlp_coarse.lin_out.weight.grad, "mlp_coarse.lin_out.bias": net.mlp_coarse.lin_out.bias.grad,
         "feats": net.encoder.feats.grad}
    if net.learn_empty:
        g["empty_feature"] = net.empty_feature.grad
    for bi, blk in enumerate(net.mlp_coarse.blocks):
        for nm, prm in (("fc_0.weight", blk.fc_0.weight), ("fc_0.bias", blk.fc_0.bias), ("fc_1.weight", blk.fc_1.weight),
                        ("fc_1.bias", blk.fc_1.bias)):
            g[f"mlp_coarse.blocks.{bi}.{nm}"] = prm.grad
    return {k: v.detach().cpu() for k, v in g.items()}, {k: v.detach().cpu() for k, v in out.items()}


GRAD_FIXTURES = [n for n in golden_names() if load_golden(n).grad]



@pytest.mark.gpu
@pytest.mark.parametrize("d_hidden", [96, 128])
def test_backward_wide_field_vs_oracle_autograd(d_hidden):
    """d_hidden 96 / 128 without blocks (the stress configuration's field width): the CUDA-core backward, 3 / 4 hidden units
    per lane, against torch autograd through the oracle."""
    wl = syn.cfg2_kitti360_eval(H=24, W=64, K=64, stereo=True, learn_empty=True, lin_out_bias=-4.0, seed=87, ray_subsample=500)
    wl.model_conf["mlp_coarse"]["d_hidden"] = d_hidden
    cpu_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in cpu_net.state_dict().items() if not k.startswith("encoder.")}
    got, _ = cuda_gradients(wl, sd)
    ref = gu.oracle_gradients(wl, sd)
    tol = gu.grad_tolerances(ref, gu.oracle_gradients64(wl, sd), GRAD_REL_TOL)
    print(d_hidden, gu.assert_grads_close(got, ref, tol, f"d_hidden={d_hidden}"), "bounds", tol)

@pytest.mark.parametrize("name", GRAD_FIXTURES)
def test_backward_matches_reference_gradient_fixture(name):
    g = load_golden(name)
    wl = pu.workload_from_golden(g)
    got, out = cuda_gradients(wl, g.sd)
    errs = {}
    for k, ref in g.grad.items():
        assert got[k].shape == ref.shape, (k, got[k].shape, ref.shape)
        assert torch.isfinite(got[k]).all(), k
        errs[k] = gu.rel_err(got[k], ref)
    print(name, errs)
    assert max(errs.values()) <= GRAD_REL_TOL, errs


def test_backward_training_shape_vs_oracle_autograd():
    """cfg3-shaped training batch: n=2 frames, 2 colour views, 2048 rays each, K=64, all four upstream gradients."""
    wl = syn.cfg3_kitti_raw_train(H=48, W=160, n=2, lin_out_bias=-4.0, seed=41, ray_subsample=2048)
    cpu_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in cpu_net.state_dict().items() if not k.startswith("encoder.")}
    got, _ = cuda_gradients(wl, sd)
    ref = gu.oracle_gradients(wl, sd)
    errs = {k: gu.rel_err(got[k], ref[k]) for k in ref}
    print(errs)
    assert max(errs.values()) <= GRAD_REL_TOL, errs


def test_backward_with_in_kernel_jitter_is_consistent():
    """Without injected u the backward must regenerate the forward's jitter from the saved seed: compare with a
    finite-difference directional derivative along lin_out.bias (a smooth direction)."""
    wl = syn.cfg2_kitti360_eval(H=24, W=80, K=64, stereo=True, learn_empty=True, lin_out_bias=-4.0, seed=42, ray_subsample=1024)
    net, wrapped = syn.build_model(wl, "cuda")
    net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    rays = wl.rays.cuda()

    def loss_at(delta, need_grad):
        with torch.no_grad():
            net.mlp_coarse.lin_out.bias.add_(delta)
        torch.manual_seed(9)                      # same seed -> same in-kernel jitter
        with torch.set_grad_enabled(need_grad):
            out = wrapped(rays, want_weights=True, want_alphas=True)["coarse"]
            l = out["depth"].sum() * 0.01 + out["rgb"].sum()
        if need_grad:
            l.backward()
        with torch.no_grad():
            net.mlp_coarse.lin_out.bias.sub_(delta)
        return l.item()

    loss_at(0.0, True)
    g = net.mlp_coarse.lin_out.bias.grad.item()
    eps = 1e-2
    fd = (loss_at(eps, False) - loss_at(-eps, False)) / (2 * eps)
    print("analytic", g, "finite difference", fd)
    assert abs(g - fd) <= 2e-2 * max(1.0, abs(fd))


def test_backward_unsupported_configs_raise():
    wl = syn.cfg4_re10k_nvs(H=24, W=36, K=64, ray_subsample=64)      # one ResnetFC block ...
    wl.model_conf["mlp_coarse"]["d_hidden"] = 64                    # ... at d_hidden = 64: inference only
    net, wrapped = syn.build_model(wl, "cuda")
    net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    with pytest.raises(NotImplementedError):
        wrapped(wl.rays.cuda())                                     # parameters require grad -> autograd path -> blocks
    with torch.no_grad():
        wrapped(wl.rays.cuda())                                     # inference is fine


@pytest.mark.parametrize("variant", ["white_linear_nocap", "white_linear_nocap_edge", "distance_code_K32", "general_forward_kernel"])
def test_backward_option_coverage_vs_oracle_autograd(variant):
    """Renderer / field options that change the backward formulas: white background (rgb depends on sum of weights),
    no hard alpha cap (last sample carries gradient), linear depth sampling + linear depth code, distance code, K=32,
    in-frustum exits (learn_empty rows), and the CUDA-core forward (precision=2) feeding the same backward.
    ``white_linear_nocap`` keeps the rays inside the frustum (well conditioned: the formulas are checked at the tight
    tolerance); ``_edge`` sends most samples behind / outside the encoder camera, where the fp32 reference itself is only
    accurate to ~1e-4 (measured against its float64 evaluation) -- bounds follow that measurement."""
    P = syn._pose
    edge = False
    if variant.startswith("white_linear_nocap"):
        edge = variant.endswith("_edge")
        mc = syn.kitti_model_conf(learn_empty=True, z_near=0.5, z_far=20.0, inv_z=False)
        rc = syn.kitti_renderer_conf(64, lindisp=False, hard_alpha_cap=False, white_bkgd=True)
        ray_pose = P(tx=0.4, tz=0.5, yaw_deg=25.0) if edge else P(tx=0.1, ty=-0.05, tz=0.2, yaw_deg=3.0)
        wl = syn.make_workload("bw_white", H=24, W=64, poses=[P(), P(tx=0.3, yaw_deg=8.0)], model_conf=mc, renderer_conf=rc,
                               ids_encoder=[0], ids_render=[0, 1], ray_poses=[ray_pose],
                               ray_subsample=700, seed=81, lin_out_bias=-2.0)
        prec = 0
    elif variant == "distance_code_K32":
        mc = syn.kitti_model_conf(learn_empty=True, code_mode="distance")
        rc = syn.kitti_renderer_conf(32)
        wl = syn.make_workload("bw_dist", H=24, W=64, poses=[P(), P(tx=0.54)], model_conf=mc, renderer_conf=rc,
                               ids_encoder=[0], ids_render=[1], ray_poses=[P(tx=0.2, tz=-0.5, yaw_deg=12.0)],
                               ray_subsample=700, seed=82, lin_out_bias=-3.0)
        prec = 0
    else:
        wl = syn.cfg2_kitti360_eval(H=24, W=64, K=64, stereo=True, learn_empty=True, lin_out_bias=-4.0, seed=83, ray_subsample=700)
        prec = 2
    cpu_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in cpu_net.state_dict().items() if not k.startswith("encoder.")}
    got, _ = cuda_gradients(wl, sd, precision=prec)
    ref = gu.oracle_gradients(wl, sd)
    # flag membership of single knife-edge samples also moves d_empty in the edge workload: floor of 1e-3 there
    tol = gu.grad_tolerances(ref, gu.oracle_gradients64(wl, sd), 1e-3 if edge else GRAD_REL_TOL)
    print(variant, gu.assert_grads_close(got, ref, tol, variant), "bounds", tol)


def test_backward_with_masked_rays_vs_oracle_autograd():
    """Rays with an all-zero upstream gradient (the loss's invalid mask): the backward skips their sample blocks; the
    gradients must equal autograd's."""
    wl = syn.cfg3_kitti_raw_train(H=48, W=160, n=2, lin_out_bias=-4.0, seed=43, ray_subsample=1024)
    cpu_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in cpu_net.state_dict().items() if not k.startswith("encoder.")}
    got, _ = cuda_gradients(wl, sd, loss_fn=gu.masked_probe_loss)
    ref = gu.oracle_gradients(wl, sd, loss_fn=gu.masked_probe_loss)
    errs = {k: gu.rel_err(got[k], ref[k]) for k in ref}
    print(errs)
    assert max(errs.values()) <= GRAD_REL_TOL, errs


def test_tensor_core_backward_matches_cuda_core_backward(monkeypatch):
    """The tcgen05 backward (d_hidden 64, no blocks, saved sigma) and the CUDA-core backward on the same training batch:
    same gradients up to accumulation order; both within tolerance of autograd through the oracle."""
    from behindthescenes_b200 import _lib
    lib = _lib.load()
    wl = syn.cfg3_kitti_raw_train(H=48, W=160, n=2, lin_out_bias=-4.0, seed=47, ray_subsample=2048)
    cpu_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in cpu_net.state_dict().items() if not k.startswith("encoder.")}
    got_tc, _ = cuda_gradients(wl, sd)
    assert lib.bts_last_bwd_kernel() == b"tc" and lib.bts_debug_tc_timeouts() == 0
    monkeypatch.setenv("BTS_BWD_SIMT", "1")
    got_si, _ = cuda_gradients(wl, sd)
    assert lib.bts_last_bwd_kernel() == b"general"
    errs = {k: gu.rel_err(got_tc[k], got_si[k]) for k in got_si}
    print(errs)
    assert max(errs.values()) <= 1e-4, errs
    ref = gu.oracle_gradients(wl, sd)
    # this batch is less well conditioned than the seed-41 one above (ReLU-mask flips): bound by the fp32-vs-fp64 reference
    tol = gu.grad_tolerances(ref, gu.oracle_gradients64(wl, sd), GRAD_REL_TOL)
    print("tolerances", tol)
    print(gu.assert_grads_close(got_tc, ref, tol, "tc backward vs oracle"))


@pytest.mark.parametrize("variant", ["K128_two_views", "K96_ragged_passes_block_mask"])
def test_tensor_core_backward_shapes_vs_oracle_autograd(variant):
    """Shapes the tensor-core backward handles beyond the K=64 training batch: K=128 (one CTA per SM), K=96 (three chunks),
    a ray count that is not a multiple of the 4 rays of a pass (idle warps in the last pass), whole passes with zero upstream
    gradient (skip path), learn_empty rows."""
    from behindthescenes_b200 import _lib
    P = syn._pose
    if variant == "K128_two_views":
        mc = syn.kitti_model_conf(learn_empty=True)
        rc = syn.kitti_renderer_conf(128)
        wl = syn.make_workload("bw_k128", H=24, W=64, poses=[P(), P(tx=0.54)], model_conf=mc, renderer_conf=rc,
                               ids_encoder=[0], ids_render=[0, 1], ray_poses=[P(tx=0.25, tz=-0.4, yaw_deg=10.0)],
                               ray_subsample=512, seed=91, lin_out_bias=-3.0)
        loss_fn = gu.probe_loss
    else:
        mc = syn.kitti_model_conf(learn_empty=False)
        rc = syn.kitti_renderer_conf(96)
        wl = syn.make_workload("bw_k96", H=24, W=64, poses=[P(), P(tx=0.54)], model_conf=mc, renderer_conf=rc,
                               ids_encoder=[0], ids_render=[1], ray_poses=[P(tx=0.3, tz=0.2, yaw_deg=-15.0)],
                               ray_subsample=333, seed=92, lin_out_bias=-3.0)
        loss_fn = gu.block_masked_probe_loss
    cpu_net, _ = syn.build_model(wl, "cpu")
    sd = {k: v.clone() for k, v in cpu_net.state_dict().items() if not k.startswith("encoder.")}
    got, _ = cuda_gradients(wl, sd, loss_fn=loss_fn)
    lib = _lib.load()
    assert lib.bts_last_bwd_kernel() == b"tc" and lib.bts_debug_tc_timeouts() == 0
    ref = gu.oracle_gradients(wl, sd, loss_fn=loss_fn)
    errs = {k: gu.rel_err(got[k], ref[k]) for k in ref}
    print(variant, errs)
    assert max(errs.values()) <= GRAD_REL_TOL, errs


class _TwoScaleEncoder(torch.nn.Module):
    """Encoder stand-in with two latents (scale 1 at half resolution), like the monodepth2 decoder's multi-scale output
    (models/common/backbones/monodepth2.py:235-237)."""

    def __init__(self, f0, f1):
        super().__init__()
        self.f0, self.f1 = torch.nn.Parameter(f0.clone()), torch.nn.Parameter(f1.clone())
        self.latent_size = f0.shape[1]
        self.scales = [0, 1]

    def forward(self, x):
        return [self.f0, self.f1]


def test_multiscale_training_backward_uses_each_forward_scene():
    """prediction_mode 'multiscale' (models/bts/trainer.py:220-242): every scale is rendered (set_scale + render) and ONE
    loss.backward() follows.  Each render's backward must scatter into the folded map of ITS scale, not the last one set."""
    import torch.nn.functional as F
    from behindthescenes_b200 import BTSNet, NeRFRenderer
    wl = syn.cfg3_kitti_raw_train(H=32, W=96, n=1, lin_out_bias=-4.0, seed=48, ray_subsample=512)
    g = torch.Generator().manual_seed(5)
    f1 = torch.randn((1, 64, 16, 48), generator=g)
    torch.manual_seed(wl.weight_seed)
    net = BTSNet(wl.model_conf, encoder=_TwoScaleEncoder(wl.feats, f1))
    with torch.no_grad():
        net.mlp_coarse.lin_out.bias.fill_(-4.0)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items() if not k.startswith("encoder.")}
    net = net.cuda()
    wrapped = NeRFRenderer.from_conf(wl.renderer_conf).bind_parallel(net, gpus=None).eval()
    net.encode(wl.images.cuda(), wl.Ks.cuda(), wl.poses_c2w.cuda(), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    rays, u = wl.rays.cuda(), wl.u.cuda()
    loss = 0.0
    for scale in (0, 1):                                   # render every scale first ...
        net.set_scale(scale)
        out = wrapped(rays, want_weights=True, want_alphas=True, u=u)["coarse"]
        loss = loss + (1.0 + scale) * gu.probe_loss(out)
    loss.backward()                                        # ... then one backward (the scene of scale 1 is current now)
    got = {"f0": net.encoder.f0.grad.cpu(), "f1": net.encoder.f1.grad.cpu(), "w_in": net.mlp_coarse.lin_in.weight.grad.cpu(),
           "w_out": net.mlp_coarse.lin_out.weight.grad.cpu()}
    # autograd through the oracle, same nearest-neighbour resize of the coarser latent (models_bts.py:126)
    f0o, f1o = wl.feats.clone().requires_grad_(True), f1.clone().requires_grad_(True)
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.startswith("code_xyz")}
    fcfg = orc_field = None
    from oracle import bts_oracle as orc
    from tests.golden_io import mlp_params
    fcfg, rcfg = orc.FieldConfig.from_model_conf(wl.model_conf), orc.RenderConfig.from_renderer_conf(wl.renderer_conf)
    total = 0.0
    for scale, f in ((0, f0o), (1, F.interpolate(f1o, (32, 96)))):
        st = orc.encode_state(wl.images, wl.Ks, wl.poses_c2w, f, wl.ids_encoder, wl.ids_render)
        out = orc.render(st, mlp_params(sdo), fcfg, rcfg, wl.rays, u=wl.u, empty_feature=sdo.get("empty_feature"),
                         want_weights=True, want_alphas=True)["coarse"]
        total = total + (1.0 + scale) * gu.probe_loss(out)
    total.backward()
    want = {"f0": f0o.grad, "f1": f1o.grad, "w_in": sdo["mlp_coarse.lin_in.weight"].grad, "w_out": sdo["mlp_coarse.lin_out.weight"].grad}
    errs = {k: gu.rel_err(got[k], want[k]) for k in want}
    print(errs)
    assert max(errs.values()) <= GRAD_REL_TOL, errs
