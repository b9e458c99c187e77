"""Patch photometric loss of the BTS training step as one fused CUDA launch each way (``bts_patch_loss_fwd/bwd``).

Host mirror of the colour term of ``ReconstructionLoss.__call__`` (``models/bts/model/loss.py:151-167``) with
``criterion: "l1+ssim"`` (``compute_errors_l1ssim``, loss.py:10-18; SSIM of ``models/common/model/layers.py:79-150``):

    rgb_loss = crit(rgb_coarse, rgb_gt).amin(-2)             # min over the rendered views
    rgb_loss = rgb_loss * (1 - invalid_coarse)                # ignore_invalid
    rgb_loss = rgb_loss.mean()

The reference runs this as ~40 elementwise / conv2d launches over ``(n*pc*nv, c, h, w)`` temporaries (and as many again in
backward); here the renderer's ``rgb`` output is consumed in the patch order it already has
(``PatchRaySampler.reconstruct`` is a view, ``ray_sampler.py:164-200``).  CUDA only, no fallback.
"""
import torch

from . import _lib


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.to(torch.float32).contiguous()


class _PatchLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, invalid_ray, invalid_mass, mass_thresh, want_errors):
        lib = _lib.load()
        n, pc, h, w, nv, c = pred.shape
        P = n * pc
        per_ray = torch.empty((n, pc, h, w), device=pred.device, dtype=torch.float32)
        argmin = torch.empty((n, pc, h, w), device=pred.device, dtype=torch.int32)
        errors = torch.empty((n, pc, h, w, nv), device=pred.device, dtype=torch.float32) if want_errors else None
        stream = torch.cuda.current_stream(pred.device).cuda_stream
        _lib.check(lib.bts_patch_loss_fwd(pred.data_ptr(), gt.data_ptr(),
                                          invalid_ray.data_ptr() if invalid_ray is not None else None,
                                          invalid_mass.data_ptr() if invalid_mass is not None else None,
                                          float(mass_thresh), P, h, w, nv, c, per_ray.data_ptr(), argmin.data_ptr(),
                                          errors.data_ptr() if errors is not None else None, stream), "bts_patch_loss_fwd")
        ctx.save_for_backward(pred, gt, invalid_ray, invalid_mass, argmin)
        ctx.mass_thresh = float(mass_thresh)
        ctx.mark_non_differentiable(argmin)
        if errors is not None:
            ctx.mark_non_differentiable(errors)
        return per_ray, argmin, errors

    @staticmethod
    def backward(ctx, g_per_ray, _g_argmin, _g_errors):
        pred, gt, invalid_ray, invalid_mass, argmin = ctx.saved_tensors
        lib = _lib.load()
        n, pc, h, w, nv, c = pred.shape
        g = _f32c(g_per_ray)
        d_pred = torch.empty_like(pred)
        stream = torch.cuda.current_stream(pred.device).cuda_stream
        _lib.check(lib.bts_patch_loss_bwd(pred.data_ptr(), gt.data_ptr(),
                                          invalid_ray.data_ptr() if invalid_ray is not None else None,
                                          invalid_mass.data_ptr() if invalid_mass is not None else None,
                                          ctx.mass_thresh, n * pc, h, w, nv, c, argmin.data_ptr(), g.data_ptr(),
                                          d_pred.data_ptr(), stream), "bts_patch_loss_bwd")
        return d_pred, None, None, None, None, None


def patch_photometric_loss(rgb, rgb_gt, invalid_ray=None, invalid_mass=None, mass_thresh=0.9, patch=None,
                           want_errors=False):
    """``rgb`` (n, pc, h, w, nv, c) -- or the renderer's raw ``(n, R, nv*c)`` with ``patch=(h, w)`` -- and ``rgb_gt``
    (n, pc, h, w, c) -- or ``(n, R, c)``.  ``invalid_ray`` (n, pc, h, w[, 1]) is the per-ray flag after the reference's
    invalid policy (loss.py:101-121); alternatively pass the renderer's ``invalid_mass`` (n, R, nv) and the
    "weight_guided" rule ``all_v(mass_v > mass_thresh)`` (loss.py:105-108) is applied inside the kernel.

    Returns ``(loss, per_ray[, errors])``: ``per_ray`` (n, pc, h, w, 1) is the masked minimum-over-views error and
    ``loss = per_ray.mean()`` (loss.py:167); ``errors`` (n, pc, h, w, nv, 1) is the un-reduced criterion (loss.py:154).
    Differentiable with respect to ``rgb``."""
    if not rgb.is_cuda:
        raise _lib.BtsCudaError("patch_photometric_loss runs on CUDA tensors only (no CPU fallback)")
    if rgb.dim() == 3:
        if patch is None:
            raise ValueError("raw (n, R, nv*c) rgb needs patch=(h, w)")
        h, w = patch
        c = rgb_gt.shape[-1]
        n, R, vc = rgb.shape
        if R % (h * w) != 0 or vc % c != 0:
            raise ValueError(f"{R} rays do not tile into {h}x{w} patches / {vc} values into {c} channels")
        rgb = rgb.view(n, R // (h * w), h, w, vc // c, c)
    n, pc, h, w, nv, c = rgb.shape
    rgb_gt = rgb_gt.reshape(n, pc, h, w, c)
    pred, gt = _f32c(rgb), _f32c(rgb_gt)
    if invalid_ray is not None:
        invalid_ray = _f32c(invalid_ray.reshape(n, pc, h, w))
        invalid_mass = None
    elif invalid_mass is not None:
        invalid_mass = _f32c(invalid_mass.reshape(n, pc, h, w, nv))
    per_ray, _, errors = _PatchLoss.apply(pred, gt, invalid_ray, invalid_mass, mass_thresh, want_errors)
    per_ray = per_ray.unsqueeze(-1)
    loss = per_ray.mean()
    if want_errors:
        return loss, per_ray, errors.unsqueeze(-1)
    return loss, per_ray
