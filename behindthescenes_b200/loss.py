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
        with _lib.on_device(pred.device):
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
        with _lib.on_device(pred.device):
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


class _PatchSmooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depth, gt, invalid_ray, invalid_mass, mass_thresh, nv):
        lib = _lib.load()
        n, pc, h, w = depth.shape
        c = gt.shape[-1]
        per_ray = torch.empty_like(depth)
        stream = torch.cuda.current_stream(depth.device).cuda_stream
        with _lib.on_device(depth.device):
            _lib.check(lib.bts_patch_smooth_fwd(depth.data_ptr(), gt.data_ptr(),
                                                invalid_ray.data_ptr() if invalid_ray is not None else None,
                                                invalid_mass.data_ptr() if invalid_mass is not None else None,
                                                float(mass_thresh), n * pc, h, w, nv, c, per_ray.data_ptr(), stream),
                       "bts_patch_smooth_fwd")
        ctx.save_for_backward(depth, gt, invalid_ray, invalid_mass)
        ctx.mass_thresh, ctx.nv = float(mass_thresh), nv
        return per_ray

    @staticmethod
    def backward(ctx, g_per_ray):
        depth, gt, invalid_ray, invalid_mass = ctx.saved_tensors
        lib = _lib.load()
        n, pc, h, w = depth.shape
        g = _f32c(g_per_ray)
        d_depth = torch.empty_like(depth)
        stream = torch.cuda.current_stream(depth.device).cuda_stream
        with _lib.on_device(depth.device):
            _lib.check(lib.bts_patch_smooth_bwd(depth.data_ptr(), gt.data_ptr(),
                                                invalid_ray.data_ptr() if invalid_ray is not None else None,
                                                invalid_mass.data_ptr() if invalid_mass is not None else None,
                                                ctx.mass_thresh, n * pc, h, w, ctx.nv, gt.shape[-1], g.data_ptr(), d_depth.data_ptr(),
                                                stream), "bts_patch_smooth_bwd")
        return d_depth, None, None, None, None, None


def patch_edge_aware_smoothness(depth, rgb_gt, invalid_ray=None, invalid_mass=None, mass_thresh=0.9):
    """Edge-aware smoothness of the rendered depth per ray (``edge_aware_smoothness``, loss.py:21-40, masked as in
    loss.py:247-256).  ``depth`` (n, pc, h, w), ``rgb_gt`` (n, pc, h, w, c); masks as in ``patch_photometric_loss``.
    Returns ``(loss_eas, per_ray)`` with ``loss_eas = per_ray.mean()``; differentiable with respect to ``depth``."""
    if not depth.is_cuda:
        raise _lib.BtsCudaError("patch_edge_aware_smoothness runs on CUDA tensors only (no CPU fallback)")
    n, pc, h, w = depth.shape
    gt = _f32c(rgb_gt.reshape(n, pc, h, w, -1))
    nv = 1
    if invalid_ray is not None:
        invalid_ray, invalid_mass = _f32c(invalid_ray.reshape(n, pc, h, w)), None
    elif invalid_mass is not None:
        nv = invalid_mass.shape[-1]
        invalid_mass = _f32c(invalid_mass.reshape(n, pc, h, w, nv))
    per_ray = _PatchSmooth.apply(_f32c(depth), gt, invalid_ray, invalid_mass, mass_thresh, nv)
    return per_ray.mean(), per_ray


def _same_tensor(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Do ``a`` and ``b`` hold the same values by construction?  The trainer aliases ``fine`` to ``coarse`` when there is
    no fine pass (trainer.py:246-247), but ``PatchRaySampler.reconstruct`` (ray_sampler.py:189-195) then replaces both
    entries with separate ``.view()`` objects of one buffer -- identity fails, storage + layout does not."""
    return a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride() and a.dtype == b.dtype)


class ReconstructionLoss:
    """Drop-in for ``models.bts.model.loss.ReconstructionLoss`` (loss.py:43-293) on the configurations the reference ships
    (``criterion: "l1+ssim"``, ``invalid_policy: weight_guided``, ``lambda_edge_aware_smoothness``; every
    ``configs/exp_*.yaml``): same constructor, same ``loss, loss_dict = criterion(data)`` contract (``utils/base_trainer.py:292``),
    same ``data`` layout (``data["coarse"][scale]`` / ``data["fine"][scale]`` dicts after ``PatchRaySampler.reconstruct``,
    ``data["rgb_gt"]``).  The photometric term and the smoothness term are one fused launch each way; the small
    regularisers (entropy, depth / alpha / surfaceness) stay torch expressions on the per-ray tensors.

    Like the reference, a ``fine`` dict that is just a copy of ``coarse`` (``models/bts/trainer.py:246-247``) is scored a second
    time with ``lambda_fine`` -- when it holds the same ``rgb`` tensor the coarse result is reused instead of recomputed.
    ``coarse["invalid_mass"]`` (n, pc, h, w, nv) from ``want_invalid_mass=True`` replaces ``invalid``/``weights`` for the
    weight-guided policy.  ``criterion`` other than ``"l1+ssim"``, automasking, median thresholding and
    ``weight_guided_diverse`` raise ``NotImplementedError``."""

    def __init__(self, config, use_automasking=False):
        self.criterion_str = config.get("criterion", "l2")
        self.invalid_policy = config.get("invalid_policy", "strict")
        if self.invalid_policy not in ("strict", "weight_guided", "weight_guided_diverse", None, "none"):
            raise ValueError(f"unknown invalid_policy {self.invalid_policy!r}")
        self.ignore_invalid = self.invalid_policy is not None and self.invalid_policy != "none"
        self.lambda_coarse = config.get("lambda_coarse", 1)
        self.lambda_fine = config.get("lambda_fine", 1)
        self.use_automasking = use_automasking
        self.lambda_entropy = config.get("lambda_entropy", 0)
        self.lambda_depth_reg = config.get("lambda_depth_reg", 0)
        self.lambda_alpha_reg = config.get("lambda_alpha_reg", 0)
        self.lambda_surfaceness_reg = config.get("lambda_surfaceness_reg", 0)
        self.lambda_edge_aware_smoothness = config.get("lambda_edge_aware_smoothness", 0)
        self.lambda_depth_smoothness = config.get("lambda_depth_smoothness", 0)
        self.median_thresholding = config.get("median_thresholding", False)
        self.alpha_reg_reduction = config.get("alpha_reg_reduction", "ray")
        self.alpha_reg_fraction = config.get("alpha_reg_fraction", 1 / 8)
        if self.alpha_reg_reduction not in ("ray", "slice"):
            raise ValueError(f"Unknown reduction for alpha regularization: {self.alpha_reg_reduction}")
        if self.criterion_str != "l1+ssim":
            raise NotImplementedError(f'criterion "{self.criterion_str}": the fused loss implements "l1+ssim" (every shipped config)')
        if use_automasking or self.median_thresholding or self.invalid_policy == "weight_guided_diverse":
            raise NotImplementedError("automasking / median thresholding / weight_guided_diverse are not implemented")
        self._weights = {}       # device-resident loss-weight vectors, one per term layout

    @staticmethod
    def get_loss_metric_names():
        return ["loss", "loss_rgb_coarse", "loss_rgb_fine", "loss_ray_entropy", "loss_depth_reg"]

    def _invalid(self, part):
        """-> (invalid_ray (n,pc,h,w) float or None, invalid_mass or None): loss.py:101-121."""
        if self.invalid_policy == "weight_guided" and "invalid_mass" in part:
            return None, part["invalid_mass"]
        inv = part["invalid"]
        if self.invalid_policy == "strict":
            ray = (inv > .5).any(-2).all(-1)
        elif self.invalid_policy == "weight_guided":
            ray = ((inv.to(torch.float32) * part["weights"].unsqueeze(-1)).sum(-2) > .9).all(-1)
        else:
            ray = torch.zeros(inv.shape[:-2], dtype=torch.bool, device=inv.device)
        return ray.to(torch.float32), None

    def __call__(self, data):
        """Every term is kept as (scalar tensor, weight in the loss, name of the logged sum, weight in that sum); the loss is
        ONE stacked dot product with a cached weight vector and the logged scalars come from ONE device->host transfer
        (the reference chains ~20 scalar kernels and calls ``.item()`` nine times, loss.py:169-291) -- the training step of
        this path is host-bound, so launches matter."""
        with torch.autograd.profiler.record_function("loss_computation"):
            n_scales = len(data["coarse"])
            coarse_0 = data["coarse"][0]
            rgb_gt = data["rgb_gt"]
            mask, keep_cache = {}, []
            inv_ray = inv_mass = None
            if self.ignore_invalid or self.lambda_entropy > 0:
                inv_ray, inv_mass = self._invalid(coarse_0)
            if self.ignore_invalid:
                mask = dict(invalid_ray=inv_ray, invalid_mass=inv_mass)

            def keep():        # (n,pc,h,w) 1 - invalid for the torch-side regularisers; built only if one of them is on
                if not keep_cache:
                    keep_cache.append(1 - (inv_ray if inv_ray is not None else (inv_mass > .9).all(-1).to(torch.float32)))
                return keep_cache[0]

            terms = []         # (tensor, loss weight, log name, log weight)
            inv_s = 1.0 / n_scales                                                                          # :269
            for scale in range(n_scales):
                coarse, fine = data["coarse"][scale], data["fine"][scale]
                rgb_loss = patch_photometric_loss(coarse["rgb"], rgb_gt, **mask)[0]                       # :151-167
                if len(fine) > 0 and _same_tensor(fine["rgb"], coarse["rgb"]):                            # :170-186, trainer.py:246-247
                    terms.append((rgb_loss, (self.lambda_coarse + self.lambda_fine) * inv_s, "loss_rgb_coarse", self.lambda_coarse))
                    terms.append((None, 0.0, "loss_rgb_fine", self.lambda_fine))                          # same tensor as the previous term
                else:
                    terms.append((rgb_loss, self.lambda_coarse * inv_s if len(fine) > 0 else inv_s, "loss_rgb_coarse", self.lambda_coarse))
                    if len(fine) > 0:
                        fmask = {}
                        if self.ignore_invalid:
                            fr, fm = self._invalid(data["fine"][0])
                            fmask = dict(invalid_ray=fr, invalid_mass=fm)
                        fine_loss = patch_photometric_loss(fine["rgb"], rgb_gt, **fmask)[0]
                        terms.append((fine_loss, self.lambda_fine * inv_s, "loss_rgb_fine", self.lambda_fine))
                depths = coarse["depth"]
                if self.lambda_depth_reg > 0:                                                             # :192-198
                    reg = (depths[:, :, 1:, :] - depths[:, :, :-1, :]).square().mean() + \
                          (depths[:, :, :, 1:] - depths[:, :, :, :-1]).square().mean()
                    terms.append((reg, self.lambda_depth_reg * inv_s, "loss_depth_reg", 1.0))
                if self.lambda_alpha_reg > 0:                                                             # :200-230
                    alphas = coarse["alphas"]
                    alpha_sum = alphas[..., :-1].sum(-1)
                    cap = torch.full_like(alpha_sum, alphas.shape[-1] * self.alpha_reg_fraction)
                    if self.ignore_invalid:
                        alpha_sum, cap = alpha_sum * keep(), cap * keep()
                    if self.alpha_reg_reduction == "ray":
                        reg = (alpha_sum - cap).clamp_min(0)
                    else:
                        reg = (alpha_sum.sum(-1) - cap.sum(-1)).clamp_min(0) / alpha_sum.shape[-1]
                    terms.append((reg.mean(), self.lambda_alpha_reg * inv_s, "loss_alpha_reg", 1.0))
                if self.lambda_surfaceness_reg > 0:                                                       # :232-245
                    alphas = coarse["alphas"]
                    sp = -torch.log(torch.exp(-alphas.abs()) + torch.exp(-(1 - alphas).abs())).mean(-1)
                    if self.ignore_invalid:
                        sp = sp * keep()
                    terms.append((sp.mean(), self.lambda_surfaceness_reg * inv_s, None, 0.0))
                if self.lambda_edge_aware_smoothness > 0:                                                 # :247-259
                    eas = patch_edge_aware_smoothness(depths, rgb_gt, **mask)[0]
                    terms.append((eas, self.lambda_edge_aware_smoothness / (2 ** scale) * inv_s, "loss_eas", 1.0))
                if self.lambda_depth_smoothness > 0:                                                      # :261-266
                    reg = (depths[..., :-1, :] - depths[..., 1:, :]).square().mean() + \
                          (depths[..., :, :-1] - depths[..., :, 1:]).square().mean()
                    terms.append((reg, self.lambda_depth_smoothness * inv_s, "loss_depth_smoothness", 1.0))
            if self.lambda_entropy > 0:                                                                   # :271-281
                import math
                alphas = coarse_0["alphas"] + 1e-5
                dens = alphas / alphas.sum(-1, keepdim=True)
                ent = -(dens * torch.log(dens)).sum(-1) / math.log2(alphas.shape[-1])
                ent = ent * keep()                                   # the reference masks with its invalid flags even when
                terms.append((ent.mean(), self.lambda_entropy, "loss_ray_entropy", 1.0))   # ignore_invalid is off (:278)
            # invalid ratio (:290): zero when the policy is "none"
            if self.ignore_invalid:
                ratio = inv_ray.mean() if inv_ray is not None else (inv_mass.amin(-1) > .9).to(torch.float32).mean()
            else:
                ratio = None

            live = [t for t in terms if t[0] is not None]
            stacked = torch.stack([t[0] for t in live])
            key = (str(stacked.device), tuple(t[1] for t in live))
            wvec = self._weights.get(key)
            if wvec is None:
                wvec = torch.tensor([t[1] for t in live], dtype=torch.float32, device=stacked.device)
                self._weights[key] = wvec
            loss = torch.dot(stacked, wvec)
        dev_vals = torch.cat((stacked.detach(), loss.detach().view(1), (ratio if ratio is not None else loss.new_zeros(())).view(1)))
        # ``sync_scalars = False`` keeps the logged values on the device (0-d tensors; ``float(v)`` when they are wanted):
        # no device->host transfer in the step, which is also what lets the whole step be captured in a CUDA graph
        host = dev_vals.tolist() if getattr(self, "sync_scalars", True) else list(dev_vals.unbind(0))
        loss_dict = {k: 0.0 for k in ("loss_rgb_coarse", "loss_rgb_fine", "loss_ray_entropy", "loss_depth_reg", "loss_alpha_reg",
                                      "loss_eas", "loss_depth_smoothness")}
        i, last = 0, 0.0
        for tensor, _, name, lw in terms:
            if tensor is not None:
                last = host[i]
                i += 1
            if name is not None:
                loss_dict[name] += last * lw
        loss_dict["loss_invalid_ratio"] = host[-1]
        loss_dict["loss"] = host[-2]
        return loss, loss_dict
