"""CPU: the drop-in classes expose the reference's object contract.  The comparisons against the
live reference run wherever /root/reference exists (the build container); the rest runs everywhere."""
import inspect

import pytest
import torch

import behindthescenes_b200 as B
from behindthescenes_b200 import synthetic as syn
from oracle import ref_import

needs_ref = pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not present")


def _ours(wl):
    return syn.build_model(wl, "cpu")


@needs_ref
@pytest.mark.parametrize("maker", [syn.cfg2_kitti360_eval, syn.cfg4_re10k_nvs])
def test_state_dict_keys_and_shapes_match_reference(maker):
    wl = maker(H=16, W=24, ray_subsample=8)
    ref = ref_import.load_reference()
    torch.manual_seed(0)
    rnet = ref_import.make_reference_net(ref, wl.model_conf, wl.feats)
    rwrapped = ref.NeRFRenderer.from_conf(wl.renderer_conf).bind_parallel(rnet, gpus=None)
    net, wrapped = _ours(wl)
    rs, os_ = rwrapped.state_dict(), wrapped.state_dict()
    assert set(rs) == set(os_), set(rs) ^ set(os_)
    for k in rs:
        assert tuple(rs[k].shape) == tuple(os_[k].shape), k
    # a reference checkpoint loads into the drop-in (utils/base_evaluator.py:50-58 uses strict=False)
    missing, unexpected = wrapped.load_state_dict(rs, strict=True)
    assert not missing and not unexpected
    for k in rs:
        assert torch.equal(wrapped.state_dict()[k], rs[k]), k


@needs_ref
def test_signatures_match_reference():
    ref = ref_import.load_reference()
    for ours, theirs, names in (
        (B.NeRFRenderer.__init__, ref.NeRFRenderer.__init__, None),
        (B.NeRFRenderer.from_conf, ref.NeRFRenderer.from_conf, None),
        (B.NeRFRenderer.bind_parallel, ref.NeRFRenderer.bind_parallel, None),
        (B.BTSNet.encode, ref.BTSNet.encode, None),
        (B.BTSNet.forward, ref.BTSNet.forward, None),
    ):
        po, pt = inspect.signature(ours).parameters, inspect.signature(theirs).parameters
        assert list(pt) == list(po)[: len(pt)], (ours, list(po), list(pt))      # ours may only append optional kwargs
        for name, prm in pt.items():
            assert po[name].default == prm.default or (prm.default is inspect._empty and po[name].default is inspect._empty), (ours, name)
        for name in list(po)[len(pt):]:
            assert po[name].default is not inspect._empty, (ours, name)
    # renderer forward: reference signature + the optional jitter injection
    po = inspect.signature(B.NeRFRenderer.forward).parameters
    pt = inspect.signature(ref.NeRFRenderer.forward).parameters
    assert list(po)[: len(pt)] == list(pt) and list(po)[len(pt):] == ["u", "extras"]   # optional extensions only


def test_live_attributes_and_wrapper_contract():
    wl = syn.cfg2_kitti360_eval(H=16, W=24, ray_subsample=8)
    net, wrapped = _ours(wl)
    r = wrapped.renderer
    assert wrapped.net is net and isinstance(r, B.NeRFRenderer)
    for a in ("n_coarse", "n_fine", "n_fine_depth", "noise_std", "depth_std", "eval_batch_size", "white_bkgd", "lindisp",
              "using_fine", "sched", "hard_alpha_cap", "iter_idx", "last_sched"):
        assert hasattr(r, a), a
    for a in ("d_min", "d_max", "learn_empty", "empty_empty", "inv_z", "color_interpolation", "code_mode", "encoder", "code_xyz",
              "flip_augmentation", "return_sample_depth", "sample_color", "mlp_coarse", "mlp_fine", "_d_in", "_d_out"):
        assert hasattr(net, a), a
    assert net.mlp_fine is None and net._d_in == 64 + 39
    assert r.using_fine is False and r.sched is None
    # callers mutate these between calls (scripts/images/gen_img_custom.py:111-112)
    r.n_coarse, r.lindisp = 32, False
    assert r.n_coarse == 32 and r.lindisp is False
    net.set_scale(1)
    assert net.get_scale() == 1
    assert net.compute_grid_transforms(None, None) is None
    # empty ray batch (nerf.py:20-24)
    rgb, depth = wrapped(torch.zeros((0, 5, 8)))
    assert rgb.shape == (0, 3) and depth.shape == (0,)
    # sched_step without a schedule is a no-op
    r.sched_step(5)
    assert int(r.iter_idx) == 0


def test_schedule_updates_sample_counts():
    r = B.NeRFRenderer(n_coarse=32, sched=[[2, 4], [48, 64], [0, 0]])
    r.sched_step(1)
    assert r.n_coarse == 32
    r.sched_step(1)
    assert r.n_coarse == 48 and int(r.last_sched) == 1
    r.sched_step(2)
    assert r.n_coarse == 64 and int(r.last_sched) == 2


def test_unsupported_configs_raise_at_construction():
    wl = syn.cfg2_kitti360_eval(H=16, W=24, ray_subsample=8)
    enc = syn.FeatureEncoder(wl.feats)
    for patch in ({"sample_color": False}, {"empty_empty": True}, {"code_mode": "bogus"}, {"mlp_coarse": {"type": "mlp"}}):
        conf = dict(wl.model_conf)
        conf.update(patch)
        with pytest.raises(NotImplementedError):
            B.BTSNet(conf, encoder=enc)
    with pytest.raises(NotImplementedError):
        B.NeRFRenderer().bind_parallel(None, gpus=[0, 1])


def test_encode_caches_like_the_reference():
    wl = syn.cfg3_kitti_raw_train(H=12, W=20, n=2, ray_subsample=8)
    net, _ = _ours(wl)
    net.encode(wl.images, wl.Ks, wl.poses_c2w, ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    assert net.grid_f_features[0].shape == (2, 1, 64, 12, 20)
    assert net.grid_f_Ks.shape == (2, 1, 3, 3) and net.grid_f_poses_w2c.shape == (2, 1, 4, 4)
    assert net.grid_c_imgs.shape == (2, 2, 3, 12, 20) and net.grid_c_Ks.shape == (2, 2, 3, 3)
    assert float(net.grid_c_imgs.min()) >= 0 and float(net.grid_c_imgs.max()) <= 1          # images * .5 + .5
    assert torch.allclose(net.grid_c_poses_w2c, torch.inverse(wl.poses_c2w)[:, wl.ids_render])
    with pytest.raises(NotImplementedError):
        net.encode(wl.images, wl.Ks, wl.poses_c2w, combine_ids=[(0, 1)])


@needs_ref
def test_reference_consumers_accept_the_output_contract():
    """The dict the drop-in returns must be consumable by the reference's own downstream code: ImageRaySampler.reconstruct
    (models/bts/model/ray_sampler.py:262-321) and distance_to_z (utils/projection_operations.py:4-16), exactly as
    models/bts/evaluator.py:68-79 chains them.  The oracle produces tensors with the drop-in's shapes/keys on the CPU
    (the GPU parity tests check that the CUDA path returns the same dict)."""
    import importlib
    from oracle import bts_oracle as orc
    from tests.golden_io import mlp_params
    ref = ref_import.load_reference()
    proj_ops = importlib.import_module("utils.projection_operations")
    H, W = 12, 20
    wl = syn.cfg2_kitti360_eval(H=H, W=W, K=32, stereo=True)
    net, wrapped = _ours(wl)
    sd = {k: v for k, v in net.state_dict().items() if not k.startswith("encoder.")}
    st = orc.encode_state(wl.images, wl.Ks, wl.poses_c2w, wl.feats, wl.ids_encoder, wl.ids_render)
    out = orc.render(st, mlp_params(sd), orc.FieldConfig.from_model_conf(wl.model_conf),
                     orc.RenderConfig.from_renderer_conf(wl.renderer_conf), wl.rays, u=wl.u,
                     empty_feature=sd.get("empty_feature"), want_weights=True, want_alphas=True)
    render_dict = dict(out)
    render_dict["fine"] = dict(render_dict["coarse"])                 # evaluator.py:70-71
    sampler = ref.ImageRaySampler(3.0, 80.0)
    rays_ref, _ = sampler.sample(wl.images * 0.5 + 0.5, wl.poses_c2w, wl.Ks)
    assert torch.allclose(rays_ref, wl.rays, atol=1e-6)               # our synthetic.gen_rays == the reference's sampler
    render_dict["rgb_gt"] = torch.zeros((1, 2 * H * W, 3))
    render_dict = sampler.reconstruct(render_dict)
    c = render_dict["coarse"]
    assert c["depth"].shape == (1, 2, H, W) and c["rgb"].shape == (1, 2, H, W, 1, 3)
    assert c["weights"].shape == (1, 2, H, W, 32) and c["invalid"].shape == (1, 2, H, W, 32, 1)
    z = proj_ops.distance_to_z(c["depth"], wl.Ks)
    assert z.shape == (1, 2, H, W) and torch.isfinite(z).all()


@needs_ref
def test_reconstruction_loss_constructor_matches_reference_class():
    """behindthescenes_b200.loss.ReconstructionLoss parses the `loss:` config block exactly like the reference class
    (models/bts/model/loss.py:44-77): same attribute names and values, same metric names, same rejection of a bad reduction."""
    from oracle import ref_import
    ref_import.load_reference()
    import importlib
    ref_cls = importlib.import_module("models.bts.model.loss").ReconstructionLoss
    from behindthescenes_b200.loss import ReconstructionLoss
    confs = [{"criterion": "l1+ssim", "invalid_policy": "weight_guided", "lambda_edge_aware_smoothness": 0.001},
             {"criterion": "l1+ssim", "invalid_policy": "strict", "lambda_fine": 0.5, "lambda_entropy": 0.1, "lambda_depth_reg": 0.01,
              "lambda_alpha_reg": 0.05, "lambda_surfaceness_reg": 0.02, "lambda_depth_smoothness": 0.03, "alpha_reg_fraction": 0.25,
              "alpha_reg_reduction": "slice"},
             {"criterion": "l1+ssim", "invalid_policy": "none"}]
    names = ["criterion_str", "invalid_policy", "ignore_invalid", "lambda_coarse", "lambda_fine", "use_automasking", "lambda_entropy",
             "lambda_depth_reg", "lambda_alpha_reg", "lambda_surfaceness_reg", "lambda_edge_aware_smoothness",
             "lambda_depth_smoothness", "median_thresholding", "alpha_reg_reduction", "alpha_reg_fraction"]
    for conf in confs:
        ours, ref = ReconstructionLoss(conf), ref_cls(conf)
        for nm in names:
            assert getattr(ours, nm) == getattr(ref, nm), (conf, nm)
    assert ReconstructionLoss.get_loss_metric_names() == ref_cls.get_loss_metric_names()
    for cls in (ReconstructionLoss, ref_cls):
        with pytest.raises(ValueError):
            cls({"criterion": "l1+ssim", "alpha_reg_reduction": "patch"})


def test_patch_rays_are_pixel_patches_of_the_ray_views():
    """``synthetic.with_patch_rays`` (the training batches of the bench and the sparse-fold tests): every 64 consecutive rays
    are one 8x8 pixel window of a ray view, row-major -- the layout ``PatchRaySampler.reconstruct`` views as (n, P, 8, 8)
    (models/bts/model/ray_sampler.py:189-195) -- and the frames draw different patches."""
    from behindthescenes_b200 import synthetic as syn
    H, W, n, P = 40, 96, 3, 5
    wl = syn.cfg3_kitti_raw_train(H=H, W=W, n=n, seed=3)
    full = wl.rays.clone().view(n, 2, H, W, 8)
    wl = syn.with_patch_rays(wl, H, W, 2, P, seed=3)
    assert wl.rays.shape == (n, P * 64, 8) and wl.u.shape == (n * P * 64, wl.renderer_conf["n_coarse"])
    flat = full.view(n, 2 * H * W, 8)
    for b in range(n):
        for i in range(P):
            patch = wl.rays[b, i * 64:(i + 1) * 64].view(8, 8, 8)
            first = (flat[b] == patch[0, 0]).all(-1).nonzero()[0, 0].item()      # (view, y, x) of the patch corner
            v, rem = divmod(first, H * W)
            y, x = divmod(rem, W)
            assert y + 8 <= H and x + 8 <= W
            assert torch.equal(patch, full[b, v, y:y + 8, x:x + 8])
    assert not torch.equal(wl.rays[0], wl.rays[1])
