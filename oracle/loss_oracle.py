"""CPU oracle of the patch photometric loss (torch restatement) -- TEST INFRASTRUCTURE ONLY, like bts_oracle.py.

Restates ``compute_errors_l1ssim`` (models/bts/model/loss.py:10-18), the SSIM it calls with
``pad_reflection=False, gaussian_average=True, comp_mode=True`` (models/common/model/layers.py:79-150) and the
reduction ``ReconstructionLoss.__call__`` applies for ``criterion: "l1+ssim"`` (loss.py:153-165): minimum over the
render views, times (1 - invalid ray), mean.  Pinned to ``tests/golden/loss/l1ssim.npz`` (produced by the unmodified
reference functions, oracle/make_goldens.py).
"""
import torch
import torch.nn.functional as F

GAUSS = ((0.0947, 0.1183, 0.0947), (0.1183, 0.1478, 0.1183), (0.0947, 0.1183, 0.0947))   # layers.py:82-85
C1, C2 = 0.01 ** 2, 0.03 ** 2                                                              # layers.py:118-119


def _gauss(x):
    k = torch.tensor(GAUSS, dtype=x.dtype, device=x.device).repeat(x.shape[1], 1, 1, 1)
    return F.conv2d(x, k, padding=0, groups=x.shape[1])                                    # layers.py:87-89


def ssim_comp(x, y):
    """(B,C,h,w) -> (B,C,h,w): clamp(1 - SSIM, 0, 1) / 2 with zero padding and the 3x3 gaussian window."""
    x, y = F.pad(x, (1, 1, 1, 1)), F.pad(y, (1, 1, 1, 1))                                  # ZeroPad2d(1), layers.py:116
    mu_x, mu_y = _gauss(x), _gauss(y)
    sx = _gauss(x ** 2) - mu_x ** 2
    sy = _gauss(y ** 2) - mu_y ** 2
    sxy = _gauss(x * y) - mu_x * mu_y
    n = (2 * mu_x * mu_y + C1) * (2 * sxy + C2)
    d = (mu_x ** 2 + mu_y ** 2 + C1) * (sx + sy + C2)
    return torch.clamp(1 - n / d, 0, 1) / 2                                                 # comp_mode, layers.py:145


def errors_l1ssim(pred, gt):
    """pred (n,pc,h,w,nv,c), gt (n,pc,h,w,1,c) -> (n,pc,h,w,nv,1)   (loss.py:10-18)."""
    n, pc, h, w, nv, c = pred.shape
    gt = gt.expand(pred.shape)
    a = pred.permute(0, 1, 4, 5, 2, 3).reshape(-1, c, h, w)
    b = gt.permute(0, 1, 4, 5, 2, 3).reshape(-1, c, h, w)
    e = 0.85 * ssim_comp(a, b).mean(1) + 0.15 * (a - b).abs().mean(1)
    return e.view(n, pc, nv, h, w).permute(0, 1, 3, 4, 2).unsqueeze(-1)


def patch_loss(pred, gt, invalid_ray=None):
    """Scalar of loss.py:153-165 (no automasking / median thresholding): mean over rays of min over views, masked."""
    e = errors_l1ssim(pred, gt.unsqueeze(-2)).amin(-2)                                     # (n,pc,h,w,1)
    if invalid_ray is not None:
        e = e * (1 - invalid_ray.to(torch.float32)).unsqueeze(-1)
    return e.mean(), e.squeeze(-1)


def edge_aware_smoothness(gt, depth):
    """gt (n,pc,h,w,3), depth (n,pc,h,w) -> (n,pc,h,w)   (loss.py:21-40; the F.interpolate there is the identity)."""
    n, pc, h, w = depth.shape
    img = gt.permute(0, 1, 4, 2, 3).reshape(-1, gt.shape[-1], h, w)
    d = 1 / depth.reshape(-1, 1, h, w).clamp(1e-3, 80)
    d = d / d.mean(dim=[2, 3], keepdim=True)
    ex = (d[:, :, :, :-1] - d[:, :, :, 1:]).abs() * torch.exp(-(img[:, :, :, :-1] - img[:, :, :, 1:]).abs().mean(1, keepdim=True))
    ey = (d[:, :, :-1, :] - d[:, :, 1:, :]).abs() * torch.exp(-(img[:, :, :-1, :] - img[:, :, 1:, :]).abs().mean(1, keepdim=True))
    return (F.pad(ex, (0, 1)) + F.pad(ey, (0, 0, 0, 1))).view(n, pc, h, w)


def smooth_loss(depth, gt, invalid_ray=None):
    """Masked mean of loss.py:247-256 (the interpolate + ceil of the mask is the identity at equal sizes)."""
    e = edge_aware_smoothness(gt, depth)
    if invalid_ray is not None:
        e = e * (1 - invalid_ray.to(torch.float32))
    return e.mean(), e
