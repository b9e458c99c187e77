"""CPU oracle for the BTS render hot path -- a torch (CPU) restatement of the reference algorithm.

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs as the checker / the CPU baseline.  Never imported by the product
package ``behindthescenes_b200`` -- the product path fails loudly without its CUDA library.

Parity pin: the reference ships no tests or golden vectors for this path, so this restatement is
pinned against the UNMODIFIED reference itself, imported in the build container (``oracle/ref_import.py``):
``oracle/make_goldens.py`` ran the real ``BTSNet``/``NeRFRenderer`` and committed inputs+outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against those vectors and
``tests/test_oracle_vs_reference.py`` re-runs the live comparison wherever ``/root/reference`` exists.

The reference path is torch ops end to end, so the restatement keeps the same numerics-defining torch calls
(``F.grid_sample(bilinear, border, align_corners=False)``, ``F.softplus``, ``torch.cumprod``, ``F.linear``) and the
same evaluation order, chunked by ``eval_batch_size`` points like ``models/common/render/nerf.py:244-268`` so it can
also serve as the CPU baseline ("port").  Each function cites the reference lines it follows
(paths relative to the reference root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

EPS = 1e-3  # models/bts/model/models_bts.py:14


@dataclass
class FieldState:
    """What ``BTSNet.encode`` caches (models/bts/model/models_bts.py:128-136)."""
    feats: torch.Tensor      # (n, Vf, C, H, W)
    Kf: torch.Tensor         # (n, Vf, 3, 3)
    Mf: torch.Tensor         # (n, Vf, 4, 4) world->cam
    imgs: torch.Tensor       # (n, Vc, Cc, Hc, Wc) in [0, 1]
    Kc: torch.Tensor         # (n, Vc, 3, 3)
    Mc: torch.Tensor         # (n, Vc, 4, 4)


@dataclass
class FieldConfig:
    z_near: float
    z_far: float
    inv_z: bool = True
    code_mode: str = "z"
    learn_empty: bool = True
    num_freqs: int = 6
    freq_factor: float = 1.5
    include_input: bool = True
    n_blocks: int = 0

    @classmethod
    def from_model_conf(cls, conf: Dict) -> "FieldConfig":
        code = conf.get("code", {})
        return cls(z_near=conf["z_near"], z_far=conf["z_far"], inv_z=conf.get("inv_z", True),
                   code_mode=conf.get("code_mode", "z"), learn_empty=conf.get("learn_empty", True),
                   num_freqs=code.get("num_freqs", 6), freq_factor=code.get("freq_factor", math.pi),
                   include_input=code.get("include_input", True),
                   n_blocks=conf["mlp_coarse"].get("n_blocks", 5))


@dataclass
class RenderConfig:
    n_coarse: int = 128
    lindisp: bool = True
    hard_alpha_cap: bool = False
    white_bkgd: bool = False
    eval_batch_size: int = 100000

    @classmethod
    def from_renderer_conf(cls, conf: Dict) -> "RenderConfig":
        return cls(n_coarse=conf.get("n_coarse", 128), lindisp=conf.get("lindisp", True),
                   hard_alpha_cap=conf.get("hard_alpha_cap", False), white_bkgd=conf.get("white_bkgd", False),
                   eval_batch_size=conf.get("eval_batch_size", 100000))


# ----------------------------------------------------------------------------------------------------------------
# encode-side glue (boundary producer)
# ----------------------------------------------------------------------------------------------------------------

def encode_state(images: torch.Tensor, Ks: torch.Tensor, poses_c2w: torch.Tensor, feats: torch.Tensor,
                 ids_encoder: Optional[List[int]], ids_render: Optional[List[int]]) -> FieldState:
    """models/bts/model/models_bts.py:65-136 minus the encoder forward (``feats`` is its scale-0 output)."""
    n, v = images.shape[:2]
    w2c = torch.inverse(poses_c2w)                                   # :66
    ie = list(range(v)) if ids_encoder is None else ids_encoder
    ir = list(range(v)) if ids_render is None else ids_render
    imgs01 = images * 0.5 + 0.5                                      # :81
    C, H, W = feats.shape[-3:]
    return FieldState(feats=feats.reshape(n, len(ie), C, H, W), Kf=Ks[:, ie], Mf=w2c[:, ie],
                      imgs=imgs01[:, ir], Kc=Ks[:, ir], Mc=w2c[:, ir])


def init_mlp(d_in: int, d_hidden: int, n_blocks: int, d_out: int = 1, seed: int = 0,
             lin_out_bias: float = 0.0) -> Dict[str, torch.Tensor]:
    """Random weights with the reference's init scheme and state-dict names
    (models/common/model/resnetfc.py:36-40,86-94): kaiming-normal fan_in, zero bias, fc_1 zero."""
    g = torch.Generator().manual_seed(seed)

    def kaiming(o, i):
        return torch.randn((o, i), generator=g) * math.sqrt(2.0 / i)

    p = {"lin_in.weight": kaiming(d_hidden, d_in), "lin_in.bias": torch.zeros(d_hidden)}
    for b in range(n_blocks):
        p[f"blocks.{b}.fc_0.weight"] = kaiming(d_hidden, d_hidden)
        p[f"blocks.{b}.fc_0.bias"] = torch.zeros(d_hidden)
        # the reference zero-initialises fc_1 (resnetfc.py:40); use small random values so the block is exercised
        p[f"blocks.{b}.fc_1.weight"] = kaiming(d_hidden, d_hidden) * 0.5
        p[f"blocks.{b}.fc_1.bias"] = torch.zeros(d_hidden)
    p["lin_out.weight"] = kaiming(d_out, d_hidden)
    p["lin_out.bias"] = torch.full((d_out,), float(lin_out_bias))
    return p


# ----------------------------------------------------------------------------------------------------------------
# field evaluation  (models/bts/model/models_bts.py:138-338)
# ----------------------------------------------------------------------------------------------------------------

def positional_encoding(x: torch.Tensor, num_freqs: int, freq_factor: float, include_input: bool) -> torch.Tensor:
    """models/common/model/code.py:12-42: out = [x, sin(f0 x), sin(f0 x + pi/2), sin(f1 x), ...] with
    f_i = freq_factor * 2^i; layout octave-major, then (sin, cos), then input dim."""
    freqs = freq_factor * 2.0 ** torch.arange(0, num_freqs)
    fr = torch.repeat_interleave(freqs, 2).view(1, -1, 1).to(x.device)   # :21-23 (registered buffers follow the module)
    ph = torch.zeros(2 * num_freqs)
    ph[1::2] = math.pi * 0.5                                         # :26-28
    ph = ph.view(1, -1, 1).to(x.device)
    rep = x.unsqueeze(1).repeat(1, num_freqs * 2, 1)                 # :37
    emb = torch.sin(torch.addcmul(ph, rep, fr)).view(x.shape[0], -1)  # :38-39
    return torch.cat((x, emb), dim=-1) if include_input else emb      # :40-41


def _project(xyz: torch.Tensor, Kmat: torch.Tensor, M: torch.Tensor):
    """models_bts.py:145-155 / :221-232.  xyz (n,P,3); Kmat (n,V,3,3); M (n,V,4,4) -> uv (n,V,P,2), z, dist, invalid."""
    h = torch.cat((xyz.unsqueeze(1), torch.ones_like(xyz[..., :1]).unsqueeze(1)), dim=-1)   # (n,1,P,4)
    cam = M[:, :, :3, :] @ h.permute(0, 1, 3, 2)                      # (n,V,3,P)           :148
    dist = torch.norm(cam, dim=-2).unsqueeze(-1)                      # (n,V,P,1)           :149
    pix = (Kmat @ cam).permute(0, 1, 3, 2)                            # (n,V,P,3)           :150
    z = pix[..., 2:3]
    uv = pix[..., :2] / z.clamp_min(EPS)                              # :154
    invalid = (z <= EPS) | (uv[..., :1] < -1) | (uv[..., :1] > 1) | (uv[..., 1:2] < -1) | (uv[..., 1:2] > 1)  # :155
    return uv, z, dist, invalid


def _bilinear_border(maps: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    """maps (n,V,C,H,W), uv (n,V,P,2) -> (n,V,P,C)   (models_bts.py:179, :234)."""
    n, V, C, H, W = maps.shape
    P = uv.shape[2]
    out = F.grid_sample(maps.reshape(n * V, C, H, W), uv.reshape(n * V, 1, P, 2), mode="bilinear",
                        padding_mode="border", align_corners=False)
    return out.view(n, V, C, P).permute(0, 1, 3, 2)


def mlp_resnetfc(x: torch.Tensor, p: Dict[str, torch.Tensor], n_blocks: int) -> torch.Tensor:
    """models/common/model/resnetfc.py:132-184 with d_latent=0 and ReLU activations; block = :53-62."""
    h = F.linear(x, p["lin_in.weight"], p["lin_in.bias"])
    for b in range(n_blocks):
        t = F.linear(torch.relu(h), p[f"blocks.{b}.fc_0.weight"], p[f"blocks.{b}.fc_0.bias"])
        h = h + F.linear(torch.relu(t), p[f"blocks.{b}.fc_1.weight"], p[f"blocks.{b}.fc_1.bias"])
    return F.linear(torch.relu(h), p["lin_out.weight"], p["lin_out.bias"])


def field_forward(st: FieldState, mlp: Dict[str, torch.Tensor], cfg: FieldConfig, xyz: torch.Tensor,
                  empty_feature: Optional[torch.Tensor] = None, only_density: bool = False):
    """``BTSNet.forward`` (models_bts.py:266-338) for the shipped configuration family: one encoder view,
    sample_color=True, no combine groups, empty_empty=False.
    xyz (n,P,3) -> rgb (n,P,3*Vc), invalid (n,P,Vc) float, sigma (n,P,1)."""
    n, P, _ = xyz.shape
    Vf = st.feats.shape[1]
    assert Vf == 1, "oracle covers the single-encoder-view path used by every shipped config"
    uv, z, dist, inv_f = _project(xyz, st.Kf, st.Mf)
    zeta = z if cfg.code_mode == "z" else dist                        # :157-171
    if cfg.inv_z:
        zeta = (1 / zeta.clamp_min(EPS) - 1 / cfg.z_far) / (1 / cfg.z_near - 1 / cfg.z_far)
    else:
        zeta = (zeta - cfg.z_near) / (cfg.z_far - cfg.z_near)
    zeta = 2 * zeta - 1
    coded = positional_encoding(torch.cat((uv, zeta), dim=-1).view(n * Vf * P, 3), cfg.num_freqs, cfg.freq_factor,
                                cfg.include_input).view(n, Vf, P, -1)  # :172
    feat = _bilinear_border(st.feats, uv)                             # :179
    if cfg.learn_empty:                                               # :181-182
        C = feat.shape[-1]
        feat = torch.where(inv_f.expand(-1, -1, -1, C), empty_feature.view(1, 1, 1, C).expand_as(feat), feat)
    x = torch.cat((feat, coded), dim=-1)                              # :184  (features first, code last)
    inv_f_views = inv_f                                               # (n,Vf,P,1)
    x = x.mean(dim=1)                                                 # :212-214 (identity for Vf=1; the reference
    inv_f = torch.any(inv_f, dim=1)                                   #  skips it when only_density, :283)  (n,P,1)
    sigma = F.softplus(mlp_resnetfc(x, mlp, cfg.n_blocks)[..., :1])   # :294-313
    Vc = st.imgs.shape[1]
    if only_density:                                                  # :335-337 -- invalid keeps the view axis:
        return torch.zeros((n, P, Vc * 3), device=sigma.device), inv_f_views.to(sigma.dtype), sigma   # (n,Vf,P,1), see :283 + :337
    uvc, _, _, inv_c = _project(xyz, st.Kc, st.Mc)                    # :221-232
    col = _bilinear_border(st.imgs, uvc)                              # (n,Vc,P,Cc)  :234
    Cc = col.shape[-1]
    rgb = col.permute(0, 2, 1, 3).reshape(n, P, Vc * Cc)              # :330  view-major
    inv_c = inv_c.permute(0, 2, 1, 3).reshape(n, P, Vc)               # :331
    invalid = (inv_c | inv_f).to(rgb.dtype)                           # :333-334
    return rgb, invalid, sigma


# ----------------------------------------------------------------------------------------------------------------
# renderer  (models/common/render/nerf.py)
# ----------------------------------------------------------------------------------------------------------------

def stratified_depths(rays: torch.Tensor, n_coarse: int, lindisp: bool, u: Optional[torch.Tensor]) -> torch.Tensor:
    """nerf.py:103-123.  rays (B,8); u (B,K) jitter in [0,1) (``torch.rand_like`` in the reference)."""
    near, far = rays[:, -2:-1], rays[:, -1:]
    step = 1.0 / n_coarse
    steps = torch.linspace(0, 1 - step, n_coarse, device=rays.device).unsqueeze(0).repeat(rays.shape[0], 1)
    if u is None:
        u = torch.rand_like(steps)
    steps = steps + u * step
    if not lindisp:
        return near * (1 - steps) + far * steps
    return 1 / (1 / near * (1 - steps) + 1 / far * steps)


def render(st: FieldState, mlp: Dict[str, torch.Tensor], fcfg: FieldConfig, rcfg: RenderConfig, rays: torch.Tensor,
           u: Optional[torch.Tensor] = None, empty_feature: Optional[torch.Tensor] = None,
           want_weights=False, want_alphas=False, want_z_samps=False, want_rgb_samps=False) -> Dict:
    """``NeRFRenderer.forward`` + ``composite`` + ``_format_outputs`` (nerf.py:210-401), coarse pass only
    (every shipped config has n_fine = 0).  rays (n,R,8) -> {"coarse": {...}} with the reference's shapes."""
    assert rays.dim() == 3
    n = rays.shape[0]
    flat = rays.reshape(-1, 8)
    z = stratified_depths(flat, rcfg.n_coarse, rcfg.lindisp, u)       # (B,K)
    B, K = z.shape
    deltas = torch.cat((z[:, 1:] - z[:, :-1], 1e10 * torch.ones_like(z[:, :1])), -1)        # :225-228
    pts = (flat[:, None, :3] + z.unsqueeze(2) * flat[:, None, 3:6]).reshape(n, -1, 3)       # :231, :244
    chunk = (rcfg.eval_batch_size - 1) // n + 1                       # :246
    rgbs, invs, sigs = [], [], []
    for part in torch.split(pts, chunk, dim=1):                       # :264-268
        r, i, s = field_forward(st, mlp, fcfg, part, empty_feature)
        rgbs.append(r), invs.append(i), sigs.append(s)
    rgbs = torch.cat(rgbs, 1).reshape(B, K, -1)
    invalid = torch.cat(invs, 1).reshape(B, K, -1)
    sigmas = torch.cat(sigs, 1).reshape(B, K)
    alphas = 1 - torch.exp(-deltas.abs() * torch.relu(sigmas))        # :283
    if rcfg.hard_alpha_cap:
        alphas[:, -1] = 1                                             # :285-286
    shifted = torch.cat((torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10), -1)           # :290-292
    T = torch.cumprod(shifted, -1)
    weights = alphas * T[:, :-1]                                      # :294
    rgb = torch.sum(weights.unsqueeze(-1) * rgbs, -2)                 # :298
    depth = torch.sum(weights * z, -1)                                # :299
    if rcfg.white_bkgd:
        rgb = rgb + 1 - weights.sum(dim=1).unsqueeze(-1)              # :301-304
    out = {"rgb": rgb.reshape(n, -1, rgb.shape[-1]), "depth": depth.reshape(n, -1),
           "invalid": invalid.reshape(n, -1, K, invalid.shape[-1])}   # :377-392
    if want_weights:
        out["weights"] = weights.reshape(n, -1, K)
    if want_alphas:
        out["alphas"] = alphas.reshape(n, -1, K)
    if want_z_samps:
        out["z_samps"] = z.reshape(n, -1, K)
    if want_rgb_samps:
        out["rgb_samps"] = rgbs.reshape(n, -1, K, rgbs.shape[-1])
    return {"coarse": out}
