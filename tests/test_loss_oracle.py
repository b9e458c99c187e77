"""CPU: the loss oracle (oracle/loss_oracle.py) against the fixture the unmodified reference loss functions produced
(tests/golden/loss/l1ssim.npz, oracle/make_goldens.py::make_loss_golden)."""
import os

import numpy as np
import torch

from oracle import loss_oracle as lo

FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss", "l1ssim.npz")


def load_loss_fixture():
    z = np.load(FIX)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_oracle_matches_reference_errors_loss_and_gradient():
    f = load_loss_fixture()
    pred = f["pred"].clone().requires_grad_(True)
    errors = lo.errors_l1ssim(pred, f["gt"].unsqueeze(-2))
    assert errors.shape == f["errors"].shape
    assert torch.equal(errors.detach(), f["errors"])           # same torch ops, same machine family: bit-exact
    loss, per_ray = lo.patch_loss(pred, f["gt"], f["invalid"])
    assert abs(loss.item() - float(f["loss"])) < 1e-8
    assert torch.allclose(per_ray.detach().unsqueeze(-1), f["per_ray"], atol=1e-8)
    loss.backward()
    assert torch.allclose(pred.grad, f["grad_pred"], atol=1e-9)


def test_fixture_covers_the_edge_cases():
    f = load_loss_fixture()
    assert (f["errors"][0, 0] == 0).all()                      # perfectly reconstructed patch: SSIM clamp at 0
    assert f["invalid"].any() and not f["invalid"].all()
    assert (f["per_ray"].squeeze(-1)[f["invalid"]] == 0).all()


RECON_FIX = os.path.join(os.path.dirname(FIX), "reconstruction_loss.npz")


def load_recon_fixture():
    import json
    z = np.load(RECON_FIX)
    out = {"in": {}, "cfg": {}}
    for k in z.files:
        head, tail = k.split("/", 1)
        if head == "in":
            out["in"][tail] = torch.from_numpy(z[k])
        else:
            v = json.loads(bytes(z[k]).decode()) if tail == "dict" else torch.from_numpy(z[k])
            out["cfg"].setdefault(head, {})[tail] = v
    return out


def test_smoothness_oracle_matches_reference_class_output():
    """loss_eas and d loss / d depth of the reference's ReconstructionLoss (weight_guided mask and unmasked)."""
    f = load_recon_fixture()
    i = f["in"]
    for name, lam, inv in (("shipped", 0.001, ((i["invalid"] * i["weights"].unsqueeze(-1)).sum(-2) > .9).all(-1)),
                           ("unmasked_slice", 0.002, None)):
        depth = i["depth"].clone().requires_grad_(True)
        loss, _ = lo.smooth_loss(depth, i["rgb_gt"], inv)
        assert abs(loss.item() - f["cfg"][name]["dict"]["loss_eas"]) < 1e-6
        (loss * lam).backward()
        assert torch.allclose(depth.grad, f["cfg"][name]["grad_depth"], atol=1e-9, rtol=1e-5)
    # the fixture exercises the clamp and a flat patch
    assert (f["cfg"]["unmasked_slice"]["grad_depth"][0, 0, :2] == 0).all()


import pytest


@pytest.mark.parametrize("name", ["shipped", "all_terms", "unmasked_slice"])
def test_reconstruction_loss_bookkeeping_on_cpu(name, monkeypatch):
    """The host logic of behindthescenes_b200.loss.ReconstructionLoss (term weights, the fine = copy-of-coarse rule, the logged
    sums, the masks handed to the fused ops) against the reference class fixture, with the two CUDA ops replaced by the
    oracle -- in this test only; the GPU suite runs the same fixtures through the real kernels."""
    from oracle.make_goldens import RECON_CONFIGS
    from behindthescenes_b200 import loss as L

    def fake_photo(rgb, rgb_gt, invalid_ray=None, invalid_mass=None, mass_thresh=0.9, **kw):
        inv = invalid_ray if invalid_ray is not None else (None if invalid_mass is None else (invalid_mass > mass_thresh).all(-1))
        loss, per_ray = lo.patch_loss(rgb, rgb_gt, None if inv is None else inv.reshape(rgb.shape[:4]))
        return loss, per_ray

    def fake_smooth(depth, rgb_gt, invalid_ray=None, invalid_mass=None, mass_thresh=0.9):
        inv = invalid_ray if invalid_ray is not None else (None if invalid_mass is None else (invalid_mass > mass_thresh).all(-1))
        return lo.smooth_loss(depth, rgb_gt, None if inv is None else inv.reshape(depth.shape))

    monkeypatch.setattr(L, "patch_photometric_loss", fake_photo)
    monkeypatch.setattr(L, "patch_edge_aware_smoothness", fake_smooth)
    f = load_recon_fixture()
    want = f["cfg"][name]
    t = {k: v.clone() for k, v in f["in"].items()}
    for k in ("rgb", "depth", "alphas"):
        t[k].requires_grad_(True)
    coarse = {k: t[k] for k in ("rgb", "depth", "alphas", "weights", "invalid")}
    loss, loss_dict = L.ReconstructionLoss(RECON_CONFIGS[name])({"coarse": [coarse], "fine": [dict(coarse)], "rgb_gt": t["rgb_gt"]})
    loss.backward()
    assert abs(loss.item() - float(want["loss"])) < 1e-6 * max(1.0, abs(float(want["loss"])))
    assert set(loss_dict) == set(want["dict"])
    for k, v in want["dict"].items():
        assert abs(loss_dict[k] - v) < 1e-6 * max(1.0, abs(v)), k
    for k in ("rgb", "depth", "alphas"):
        ref = want["grad_" + k]
        got = t[k].grad if t[k].grad is not None else torch.zeros_like(ref)
        assert torch.allclose(got, ref, rtol=1e-4, atol=1e-9 + 1e-6 * ref.abs().max().item()), k
    if name == "shipped":       # the renderer's invalid_mass output may replace invalid + weights
        mass = (f["in"]["invalid"] * f["in"]["weights"].unsqueeze(-1)).sum(-2)
        c2 = {"rgb": f["in"]["rgb"], "depth": f["in"]["depth"], "invalid_mass": mass}
        l2, d2 = L.ReconstructionLoss(RECON_CONFIGS[name])({"coarse": [c2], "fine": [dict(c2)], "rgb_gt": f["in"]["rgb_gt"]})
        assert abs(l2.item() - float(want["loss"])) < 1e-6 and abs(d2["loss_invalid_ratio"] - want["dict"]["loss_invalid_ratio"]) < 1e-7
