"""GPU parity of the fused patch photometric loss (bts_patch_loss_fwd/bwd through behindthescenes_b200.loss) against
the reference fixture and against the CPU oracle with torch autograd on larger seeded inputs.

Tolerances: the SSIM ratio amplifies fp32 summation-order differences where the window variance is ~0 (C2 = 9e-4 is the
only thing in the denominator), so per-ray values are compared at 2e-5 absolute (values are O(0.1)), the mean at 2e-6, the
gradient at 1e-3 relative in the max norm."""
import pytest
import torch

from tests.test_loss_oracle import load_loss_fixture

pytestmark = pytest.mark.gpu

VAL_TOL, MEAN_TOL, GRAD_REL = 2e-5, 2e-6, 1e-3


def _run(pred, gt, **kw):
    from behindthescenes_b200.loss import patch_photometric_loss
    dev = torch.device("cuda")
    p = pred.to(dev).clone().requires_grad_(True)
    out = patch_photometric_loss(p, gt.to(dev), **{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in kw.items()})
    out[0].backward()
    return [o.detach().cpu() for o in out], p.grad.cpu()


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def test_fixture_from_reference():
    f = load_loss_fixture()
    (loss, per_ray, errors), grad = _run(f["pred"], f["gt"], invalid_ray=f["invalid"], want_errors=True)
    assert errors.shape == f["errors"].shape and per_ray.shape == f["per_ray"].shape
    assert (errors - f["errors"]).abs().max().item() < VAL_TOL
    assert (per_ray - f["per_ray"]).abs().max().item() < VAL_TOL
    assert abs(loss.item() - float(f["loss"])) < MEAN_TOL
    assert _rel(grad, f["grad_pred"]) < GRAD_REL
    print(f"fixture: errors {(errors - f['errors']).abs().max().item():.2e} loss {abs(loss.item() - float(f['loss'])):.2e} "
          f"grad rel {_rel(grad, f['grad_pred']):.2e}")


@pytest.mark.parametrize("shape", [(4, 64, 8, 8, 3, 3), (1, 5, 4, 6, 1, 1), (2, 3, 16, 16, 2, 4), (1, 1, 1, 1, 2, 3)])
def test_against_oracle_autograd(shape):
    from oracle import loss_oracle as lo
    n, pc, h, w, nv, c = shape
    g = torch.Generator().manual_seed(sum(shape))
    gt = torch.rand((n, pc, h, w, c), generator=g)
    pred = (gt.unsqueeze(-2) + 0.2 * torch.randn(shape, generator=g)).clamp(0, 1)
    invalid = torch.rand((n, pc, h, w), generator=g) > 0.7
    ref = pred.clone().requires_grad_(True)
    loss_o, per_ray_o = lo.patch_loss(ref, gt, invalid)
    loss_o.backward()
    (loss, per_ray), grad = _run(pred, gt, invalid_ray=invalid)
    assert (per_ray.squeeze(-1) - per_ray_o.detach()).abs().max().item() < VAL_TOL
    assert abs(loss.item() - loss_o.item()) < MEAN_TOL
    assert _rel(grad, ref.grad) < GRAD_REL


def test_weight_guided_mask_in_kernel_and_raw_layout():
    """invalid_mass + threshold inside the kernel == the reference's weight_guided rule applied outside (loss.py:105-108);
    the renderer's raw (n, R, nv*c) layout with patch=(h, w) == the reconstructed 6-d layout."""
    n, pc, h, w, nv, c = 2, 8, 8, 8, 2, 3
    g = torch.Generator().manual_seed(5)
    gt = torch.rand((n, pc, h, w, c), generator=g)
    pred = torch.rand((n, pc, h, w, nv, c), generator=g)
    mass = torch.rand((n, pc * h * w, nv), generator=g)
    mass[:, ::3] = 0.95
    invalid = (mass > 0.9).all(-1)
    (l1, m1), g1 = _run(pred, gt, invalid_ray=invalid.view(n, pc, h, w, 1))
    (l2, m2), g2 = _run(pred.view(n, pc * h * w, nv * c), gt.view(n, pc * h * w, c), invalid_mass=mass, patch=(h, w))
    assert invalid.any() and torch.equal(m1, m2) and l1.item() == l2.item()
    assert torch.equal(g1.view_as(g2), g2)


def test_render_then_loss_backward_chain():
    """rgb of the fused render -> fused loss -> bts_patch_loss_bwd -> bts_render_bwd: one graph, gradients reach the MLP;
    the loss of the rendered colours equals the oracle loss of the same colours."""
    from behindthescenes_b200 import synthetic as syn
    from oracle import loss_oracle as lo
    from behindthescenes_b200.loss import patch_photometric_loss
    dev = torch.device("cuda")
    wl = syn.cfg3_kitti_raw_train(H=48, W=160, n=2, lin_out_bias=-4.0, seed=3, ray_subsample=256)
    net, wrapped = syn.build_model(wl, dev)
    net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    out = wrapped(wl.rays.to(dev), u=wl.u.to(dev), want_weights=True, want_invalid=False, want_invalid_mass=True)["coarse"]
    assert "invalid" not in out and out["invalid_mass"].shape == (2, 256, 2) and not out["invalid_mass"].requires_grad
    gt = torch.rand((2, 256, 3), generator=torch.Generator().manual_seed(1)).to(dev)
    loss, per_ray = patch_photometric_loss(out["rgb"], gt, invalid_mass=out["invalid_mass"], patch=(8, 8))
    loss.backward()
    gw = net.mlp_coarse.lin_in.weight.grad
    assert gw is not None and torch.isfinite(gw).all() and gw.abs().max() > 0
    inv = (out["invalid_mass"] > 0.9).all(-1).view(2, 4, 8, 8).cpu()
    loss_o, _ = lo.patch_loss(out["rgb"].detach().cpu().view(2, 4, 8, 8, 2, 3), gt.cpu().view(2, 4, 8, 8, 3), inv)
    assert abs(loss.item() - loss_o.item()) < MEAN_TOL


def test_errors_are_loud():
    from behindthescenes_b200 import _lib
    from behindthescenes_b200.loss import patch_photometric_loss
    with pytest.raises(_lib.BtsCudaError):
        patch_photometric_loss(torch.zeros(1, 1, 8, 8, 1, 3), torch.zeros(1, 1, 8, 8, 3))
    dev = torch.device("cuda")
    with pytest.raises(NotImplementedError):
        patch_photometric_loss(torch.zeros(1, 1, 17, 8, 1, 3, device=dev), torch.zeros(1, 1, 17, 8, 3, device=dev))
    with pytest.raises(ValueError):
        patch_photometric_loss(torch.zeros(1, 100, 6, device=dev), torch.zeros(1, 100, 3, device=dev), patch=(8, 8))


@pytest.mark.parametrize("shape", [(4, 64, 8, 8), (1, 3, 5, 7), (2, 2, 16, 16), (1, 1, 1, 1)])
def test_edge_aware_smoothness_against_oracle_autograd(shape):
    from oracle import loss_oracle as lo
    from behindthescenes_b200.loss import patch_edge_aware_smoothness
    n, pc, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 1)
    gt = torch.rand((n, pc, h, w, 3), generator=g)
    depth = 2 + 90 * torch.rand(shape, generator=g)              # some beyond the 80 m clamp
    depth[0, 0].fill_(7.0)                                       # a flat patch
    invalid = torch.rand(shape, generator=g) > 0.7
    ref = depth.clone().requires_grad_(True)
    loss_o, per_ray_o = lo.smooth_loss(ref, gt, invalid)
    loss_o.backward()
    dev = torch.device("cuda")
    tol = 1e-5 * max(1.0, per_ray_o.max().item())
    d = depth.to(dev).requires_grad_(True)
    loss, per_ray = patch_edge_aware_smoothness(d, gt.to(dev), invalid_ray=invalid.to(dev))
    loss.backward()
    got, want = per_ray.detach().cpu(), per_ray_o.detach()
    err = (got - want).abs()
    if err.max().item() >= tol:          # diagnostics for the r01 transient (never reproduced since: see the stability test below)
        worst = err.flatten().argmax().item()
        off = (err > tol).flatten().nonzero().flatten()
        pytest.fail(f"worst idx {worst} (patch {worst // (h * w)}, px {worst % (h * w)}): gpu {got.flatten()[worst].item():.7f} "
                    f"ref {want.flatten()[worst].item():.7f}; {off.numel()} elements off in patches "
                    f"{sorted(set((off // (h * w)).tolist()))[:8]}; ratio gpu/ref there {(got.flatten()[off] / want.flatten()[off]).tolist()[:8]}")
    assert abs(loss.item() - loss_o.item()) < 1e-5 * max(1.0, abs(loss_o.item()))
    assert _rel(d.grad.cpu(), ref.grad) < GRAD_REL or ref.grad.abs().max() == 0


@pytest.mark.parametrize("name", ["shipped", "all_terms", "unmasked_slice"])
def test_reconstruction_loss_dropin_vs_reference_fixture(name):
    """behindthescenes_b200.loss.ReconstructionLoss vs the unmodified reference class on the same data dict: loss,
    every logged scalar, and the gradients that reach the renderer outputs."""
    from oracle.make_goldens import RECON_CONFIGS
    from tests.test_loss_oracle import load_recon_fixture
    from behindthescenes_b200.loss import ReconstructionLoss
    f = load_recon_fixture()
    want = f["cfg"][name]
    dev = torch.device("cuda")
    t = {k: v.to(dev) for k, v in f["in"].items()}
    for k in ("rgb", "depth", "alphas"):
        t[k].requires_grad_(True)
    coarse = {k: t[k] for k in ("rgb", "depth", "alphas", "weights", "invalid")}
    data = {"coarse": [coarse], "fine": [dict(coarse)], "rgb_gt": t["rgb_gt"]}
    crit = ReconstructionLoss(RECON_CONFIGS[name])
    assert crit.get_loss_metric_names() == ["loss", "loss_rgb_coarse", "loss_rgb_fine", "loss_ray_entropy", "loss_depth_reg"]
    loss, loss_dict = crit(data)
    loss.backward()
    assert abs(loss.item() - float(want["loss"])) < 2e-6 * max(1.0, abs(float(want["loss"])))
    assert set(loss_dict) == set(want["dict"])
    for k, v in want["dict"].items():
        assert abs(loss_dict[k] - v) < 2e-6 * max(1.0, abs(v)), k
    for k in ("rgb", "depth", "alphas"):
        got = t[k].grad.cpu() if t[k].grad is not None else torch.zeros_like(f["in"][k])
        ref = want["grad_" + k]
        assert _rel(got, ref) < GRAD_REL or ref.abs().max() == 0, k


def test_reconstruction_loss_accepts_invalid_mass_instead_of_flags():
    """weight_guided with coarse["invalid_mass"] (the renderer's want_invalid_mass output) == with invalid + weights."""
    from oracle.make_goldens import RECON_CONFIGS
    from tests.test_loss_oracle import load_recon_fixture
    from behindthescenes_b200.loss import ReconstructionLoss
    f = load_recon_fixture()
    dev = torch.device("cuda")
    t = {k: v.to(dev) for k, v in f["in"].items()}
    mass = (t["invalid"] * t["weights"].unsqueeze(-1)).sum(-2)
    coarse = {"rgb": t["rgb"], "depth": t["depth"], "invalid_mass": mass}
    loss, d = ReconstructionLoss(RECON_CONFIGS["shipped"])({"coarse": [coarse], "fine": [dict(coarse)], "rgb_gt": t["rgb_gt"]})
    want = f["cfg"]["shipped"]
    assert abs(loss.item() - float(want["loss"])) < 2e-6 and abs(d["loss_invalid_ratio"] - want["dict"]["loss_invalid_ratio"]) < 1e-7


def test_reconstruction_loss_unsupported_options_raise():
    from behindthescenes_b200.loss import ReconstructionLoss
    for conf in ({"criterion": "l2"}, {"criterion": "l1+ssim", "median_thresholding": True},
                 {"criterion": "l1+ssim", "invalid_policy": "weight_guided_diverse"}):
        with pytest.raises(NotImplementedError):
            ReconstructionLoss(conf)
    with pytest.raises(NotImplementedError):
        ReconstructionLoss({"criterion": "l1+ssim"}, use_automasking=True)
    with pytest.raises(ValueError):
        ReconstructionLoss({"criterion": "l1+ssim", "alpha_reg_reduction": "patch"})


def test_smoothness_is_bit_stable_between_render_launches():
    """r01 saw one patch_smooth launch a few 1e-5 off right after render tests of the same process (twice in ~25 runs).
    200 forward + backward launches interleaved with renders of both kernel families, every one compared bit for bit with
    the first: the kernels are deterministic (compute-sanitizer memcheck / racecheck / initcheck / synccheck are clean on
    them, profiles/r02_sanitizer_loss.md)."""
    from oracle import loss_oracle as lo
    from behindthescenes_b200 import synthetic as syn
    from behindthescenes_b200.loss import patch_edge_aware_smoothness
    dev = torch.device("cuda")
    shape = (4, 64, 8, 8)
    n, pc, h, w = shape
    g = torch.Generator().manual_seed(sum(shape) + 1)
    gt = torch.rand((n, pc, h, w, 3), generator=g)
    depth = 2 + 90 * torch.rand(shape, generator=g)
    depth[0, 0].fill_(7.0)
    invalid = torch.rand(shape, generator=g) > 0.7
    _, ref = lo.smooth_loss(depth, gt, invalid)
    wl = syn.cfg2_kitti360_eval(H=48, W=160, stereo=True, lin_out_bias=-4.0, seed=31, ray_subsample=8000)
    net, wrapped = syn.build_model(wl, dev)
    with torch.no_grad():
        net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    rays = wl.rays.to(dev)
    first = None
    for it in range(200):
        with torch.no_grad():
            wrapped.renderer.band = bool(it & 1)
            wrapped(rays, want_weights=True, want_alphas=True)
        d = depth.to(dev).requires_grad_(True)
        loss, per_ray = patch_edge_aware_smoothness(d, gt.to(dev), invalid_ray=invalid.to(dev))
        loss.backward()
        got, grad = per_ray.detach().cpu(), d.grad.cpu()
        if first is None:
            first = (got.clone(), grad.clone())
            assert (got - ref).abs().max().item() < 1e-5 * max(1.0, ref.max().item())
        assert torch.equal(got, first[0]) and torch.equal(grad, first[1]), f"launch {it} differs from launch 0"
