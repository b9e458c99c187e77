"""Gradient helpers shared by the CPU (oracle vs reference fixtures) and GPU (CUDA vs oracle / fixtures) tests."""
import torch

from oracle import bts_oracle as orc
from tests.golden_io import mlp_params

GRAD_KEYS = ("mlp_coarse.lin_in.weight", "mlp_coarse.lin_in.bias", "mlp_coarse.lin_out.weight", "mlp_coarse.lin_out.bias",
             "empty_feature", "feats")


def probe_loss(coarse):
    """Same scalar as oracle/make_goldens.py::probe_loss (the gradient fixtures were produced with it)."""
    total = 0.0
    for name, phase, scale in (("rgb", 0.1, 1.0), ("depth", 0.2, 0.1), ("weights", 0.3, 1.0), ("alphas", 0.4, 1.0)):
        t = coarse[name]
        coef = torch.sin(0.37 * torch.arange(t.numel(), dtype=torch.float32, device=t.device) + phase).view_as(t)
        total = total + scale * (t * coef).sum()
    return total


def masked_probe_loss(coarse):
    """probe_loss with two thirds of the rays masked out (zero upstream gradient on every output of those rays), the
    situation the invalid-ray mask of the training loss creates (loss.py:160-161)."""
    R = coarse["depth"].shape[1]
    keep = (torch.arange(R, device=coarse["depth"].device) % 3 == 0).to(torch.float32).view(1, R)
    masked = {k: coarse[k] * keep.view((1, R) + (1,) * (coarse[k].dim() - 2)) for k in ("rgb", "depth", "weights", "alphas")}
    return probe_loss(masked)


def block_masked_probe_loss(coarse):
    """probe_loss with runs of 8 consecutive rays masked out (16 of every 24): whole 4-ray passes of the tensor-core
    backward see zero upstream gradients and take its skip path."""
    R = coarse["depth"].shape[1]
    keep = ((torch.arange(R, device=coarse["depth"].device) // 8) % 3 == 0).to(torch.float32).view(1, R)
    masked = {k: coarse[k] * keep.view((1, R) + (1,) * (coarse[k].dim() - 2)) for k in ("rgb", "depth", "weights", "alphas")}
    return probe_loss(masked)


def oracle_gradients(wl, sd, loss_fn=probe_loss):
    """torch autograd through the oracle port (CPU)."""
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point() and not k.startswith("code_xyz")}
    feats = wl.feats.clone().requires_grad_(True)
    st = orc.encode_state(wl.images, wl.Ks, wl.poses_c2w, feats, wl.ids_encoder, wl.ids_render)
    fcfg = orc.FieldConfig.from_model_conf(wl.model_conf)
    rcfg = orc.RenderConfig.from_renderer_conf(wl.renderer_conf)
    out = orc.render(st, mlp_params(sd), fcfg, rcfg, wl.rays, u=wl.u, empty_feature=sd.get("empty_feature"),
                     want_weights=True, want_alphas=True)["coarse"]
    loss_fn(out).backward()
    g = {k: sd[k].grad for k in sd if k != "feats" and sd[k].grad is not None}
    g["feats"] = feats.grad
    return g


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def oracle_gradients64(wl, sd, loss_fn=probe_loss):
    """The same autograd pass in float64 (inputs, weights and every intermediate): the yardstick for how well the fp32
    reference gradient is conditioned on a given workload."""
    import dataclasses
    to64 = lambda t: t.double() if torch.is_tensor(t) and t.is_floating_point() else t  # noqa: E731
    wl64 = dataclasses.replace(wl, **{f.name: to64(getattr(wl, f.name)) for f in dataclasses.fields(wl)})
    return oracle_gradients(wl64, {k: to64(v) for k, v in sd.items()}, loss_fn=loss_fn)


def grad_tolerances(ref32, ref64, base, factor=4.0):
    """Per-tensor bound on the relative L2 error of a gradient: ``base`` where the fp32 reference itself is accurate, and
    ``factor`` x (fp32 reference vs fp64 reference) where it is not -- ReLU-mask and frustum-flag flips of single samples
    under ulp-level forward differences move an ill-conditioned gradient sum by 1e-4 between two correct fp32 evaluations."""
    return {k: max(base, factor * rel_err(ref32[k].double(), ref64[k])) for k in ref32}


def assert_grads_close(got, ref32, tol, label=""):
    errs = {k: rel_err(got[k], ref32[k]) for k in ref32}
    bad = {k: (errs[k], tol[k] if isinstance(tol, dict) else tol) for k in errs if errs[k] > (tol[k] if isinstance(tol, dict) else tol)}
    assert not bad, (label, bad, errs)
    return errs
