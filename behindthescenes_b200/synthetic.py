"""Deterministic synthetic frustums for parity tests and ``bench.py``.

No dataset exists in the container or on the GPU box, so every workload is generated from a seed on the CPU
(so the CPU oracle and the GPU path see identical bits) following the reference's data conventions:

* intrinsics normalised to [-1, 1] image coordinates (``datasets/kitti_raw/kitti_raw_dataset.py:173-185``),
  KITTI-Raw values from ``scripts/images/gen_img_custom.py:72-77``, RE10K from ``:90-95``;
* poses are camera-to-world, first view = identity (``models/bts/evaluator.py:53-55``);
* images in [-1, 1] (``kitti_raw_dataset.py:253``);
* rays are ``[origin(3), direction(3), near, far]`` built like ``models/common/util/util.py:244-273``
  (pixel centres on ``linspace(-1, 1, W)``, unit directions).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

KITTI_RAW_K = ((1.1619, 0.0, -0.0184), (0.0, 3.8482, -0.0781), (0.0, 0.0, 1.0))
RE10K_K = ((1.0056, 0.0, 0.0), (0.0, 1.7877, 0.0), (0.0, 0.0, 1.0))


def _pose(tx=0.0, ty=0.0, tz=0.0, yaw_deg=0.0, pitch_deg=0.0) -> torch.Tensor:
    """cam->world 4x4: rotation about y (yaw) then x (pitch), translation (tx,ty,tz)."""
    a, b = math.radians(yaw_deg), math.radians(pitch_deg)
    ry = torch.tensor([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]], dtype=torch.float64)
    rx = torch.tensor([[1, 0, 0], [0, math.cos(b), -math.sin(b)], [0, math.sin(b), math.cos(b)]], dtype=torch.float64)
    m = torch.eye(4, dtype=torch.float64)
    m[:3, :3] = ry @ rx
    m[:3, 3] = torch.tensor([tx, ty, tz], dtype=torch.float64)
    return m.to(torch.float32)


def gen_rays(poses_c2w: torch.Tensor, width: int, height: int, z_near: float, z_far: float,
             focal: torch.Tensor, c: torch.Tensor, norm_dir: bool = True) -> torch.Tensor:
    """(v,4,4) poses -> (v,H,W,8) rays.  Same construction as ``models/common/util/util.py:113-149,244-273``:
    d = R * normalize([(x-cx)/fx, (y-cy)/fy, 1]) with x,y on linspace(-1,1,.); o = t."""
    v = poses_c2w.shape[0]
    dev = poses_c2w.device
    xs = torch.linspace(-1, 1, width, dtype=torch.float32, device=dev)
    ys = torch.linspace(-1, 1, height, dtype=torch.float32, device=dev)
    gx = xs.view(1, 1, width).expand(v, height, width)
    gy = ys.view(1, height, 1).expand(v, height, width)
    xy = torch.stack((gx, gy), dim=-1)
    xy = (xy - c.view(v, 1, 1, 2)) / focal.view(v, 1, 1, 2)
    dirs = torch.cat((xy, torch.ones_like(gx).unsqueeze(-1)), dim=-1)
    if norm_dir:
        dirs = dirs / torch.norm(dirs, dim=-1, keepdim=True)
    dirs_w = torch.matmul(poses_c2w[:, None, None, :3, :3], dirs.unsqueeze(-1))[..., 0]
    orig = poses_c2w[:, None, None, :3, 3].expand(-1, height, width, -1)
    near = torch.full((v, height, width, 1), float(z_near), dtype=torch.float32, device=dev)
    far = torch.full((v, height, width, 1), float(z_far), dtype=torch.float32, device=dev)
    return torch.cat((orig, dirs_w, near, far), dim=-1)


@dataclass
class Workload:
    """One synthetic frustum: everything ``encode`` + ``render`` need, all CPU fp32 tensors."""
    name: str
    model_conf: Dict
    renderer_conf: Dict
    images: torch.Tensor            # (n, v, 3, H, W) in [-1, 1]
    Ks: torch.Tensor                # (n, v, 3, 3) normalised
    poses_c2w: torch.Tensor         # (n, v, 4, 4)
    feats: torch.Tensor             # (n * len(ids_encoder), C, H, W): what the encoder would return at scale 0
    ids_encoder: List[int]
    ids_render: List[int]
    rays: torch.Tensor              # (n, R, 8)
    u: torch.Tensor                 # (n * R, K) stratified jitter in [0,1)
    lin_out_bias: float = 0.0       # SURVEY §8d fixture caveat (i): -4 spreads ray termination
    weight_seed: int = 0
    extra: Dict = field(default_factory=dict)


def kitti_model_conf(learn_empty: bool = True, d_hidden: int = 64, n_blocks: int = 0, d_out: int = 64,
                     z_near: float = 3.0, z_far: float = 80.0, inv_z: bool = True, code_mode: str = "z") -> Dict:
    """model_conf of ``configs/eval_depth.yaml`` / ``configs/exp_kitti_raw.yaml`` (the keys the hot path reads)."""
    conf = {
        "arch": "BTSNet",
        "code": {"num_freqs": 6, "freq_factor": 1.5, "include_input": True},
        "encoder": {"type": "monodepth2", "d_out": d_out},
        "mlp_coarse": {"type": "resnet", "n_blocks": n_blocks, "d_hidden": d_hidden},
        "mlp_fine": {"type": "empty", "n_blocks": 1, "d_hidden": 128},
        "z_near": z_near, "z_far": z_far, "inv_z": inv_z, "code_mode": code_mode,
        "learn_empty": learn_empty,
    }
    return conf


def kitti_renderer_conf(n_coarse: int = 64, lindisp: bool = True, hard_alpha_cap: bool = True,
                        white_bkgd: bool = False) -> Dict:
    return {"n_coarse": n_coarse, "n_fine": 0, "n_fine_depth": 0, "depth_std": 1.0, "sched": [],
            "white_bkgd": white_bkgd, "lindisp": lindisp, "hard_alpha_cap": hard_alpha_cap}


def _focal_center(K3: torch.Tensor):
    return torch.stack((K3[..., 0, 0], K3[..., 1, 1]), -1), torch.stack((K3[..., 0, 2], K3[..., 1, 2]), -1)


def make_workload(name: str, *, H: int, W: int, n: int = 1, poses: List[torch.Tensor], K3=KITTI_RAW_K,
                  model_conf: Dict, renderer_conf: Dict, ids_encoder: List[int], ids_render: List[int],
                  ray_views: Optional[List[int]] = None, ray_poses: Optional[List[torch.Tensor]] = None,
                  ray_subsample: Optional[int] = None, ray_near_far=None, seed: int = 0,
                  lin_out_bias: float = 0.0, norm_dir: bool = True) -> Workload:
    g = torch.Generator().manual_seed(seed)
    v = len(poses)
    C = model_conf["encoder"]["d_out"]
    images = torch.rand((n, v, 3, H, W), generator=g) * 2 - 1
    feats = torch.randn((n * len(ids_encoder), C, H, W), generator=g)
    Kt = torch.tensor(K3, dtype=torch.float32)
    Ks = Kt.view(1, 1, 3, 3).expand(n, v, 3, 3).contiguous()
    P = torch.stack(poses, 0).view(1, v, 4, 4).expand(n, v, 4, 4).contiguous()
    zn, zf = (model_conf["z_near"], model_conf["z_far"]) if ray_near_far is None else ray_near_far
    if ray_poses is None:
        ray_views = list(range(v)) if ray_views is None else ray_views
        rp = torch.stack([poses[i] for i in ray_views], 0)
    else:
        rp = torch.stack(ray_poses, 0)
    f, c = _focal_center(Kt.view(1, 3, 3).expand(rp.shape[0], 3, 3))
    rays1 = gen_rays(rp, W, H, zn, zf, f, c, norm_dir=norm_dir).reshape(-1, 8)
    if ray_subsample is not None and ray_subsample < rays1.shape[0]:
        idx = torch.randperm(rays1.shape[0], generator=g)[:ray_subsample].sort().values
        rays1 = rays1[idx]
    rays = rays1.view(1, -1, 8).expand(n, -1, 8).contiguous()
    K = renderer_conf["n_coarse"]
    u = torch.rand((rays.shape[0] * rays.shape[1], K), generator=torch.Generator().manual_seed(seed + 7))
    return Workload(name=name, model_conf=model_conf, renderer_conf=renderer_conf, images=images, Ks=Ks,
                    poses_c2w=P, feats=feats, ids_encoder=ids_encoder, ids_render=ids_render, rays=rays, u=u,
                    lin_out_bias=lin_out_bias, weight_seed=seed)


# --------------------------------------------------------------------------------------------------------------------
# The BASELINE.json configs.  ``scale`` shrinks H, W (parity fixtures); 1.0 = the named shape.
# --------------------------------------------------------------------------------------------------------------------

def cfg2_kitti360_eval(H=192, W=640, K=64, stereo=True, learn_empty=True, lin_out_bias=0.0, seed=0,
                       ray_subsample=None) -> Workload:
    """BASELINE config 2 (and config 1 with K=32): eval_depth, encoder view 0, rays from view 0 (+ stereo view)."""
    poses = [_pose(), _pose(tx=0.54)] if stereo else [_pose()]
    return make_workload(f"cfg2_{H}x{W}_K{K}", H=H, W=W, poses=poses, model_conf=kitti_model_conf(learn_empty),
                         renderer_conf=kitti_renderer_conf(K), ids_encoder=[0], ids_render=[0],
                         ray_views=list(range(len(poses))), seed=seed, lin_out_bias=lin_out_bias,
                         ray_subsample=ray_subsample)


def cfg3_kitti_raw_train(H=370, W=1226, K=64, n=1, seed=0, lin_out_bias=0.0, ray_subsample=None) -> Workload:
    """BASELINE config 3: 4 views, encoder view 0, Vc=2 render views, rays from the 2 loss views."""
    poses = [_pose(), _pose(tx=0.54), _pose(tz=-1.0, yaw_deg=1.0), _pose(tx=0.54, tz=-1.0, yaw_deg=1.0)]
    return make_workload(f"cfg3_{H}x{W}_K{K}", H=H, W=W, n=n, poses=poses,
                         model_conf=kitti_model_conf(learn_empty=False), renderer_conf=kitti_renderer_conf(K),
                         ids_encoder=[0], ids_render=[1, 2], ray_views=[0, 3], seed=seed,
                         lin_out_bias=lin_out_bias, ray_subsample=ray_subsample)


def with_patch_rays(wl: Workload, H: int, W: int, n_views: int, patches: int, seed: int = 0, patch: int = 8) -> Workload:
    """Replace the workload's rays by ``patches`` random ``patch`` x ``patch`` pixel patches per frame, drawn from its
    ``n_views`` ray views -- what ``PatchRaySampler.sample`` hands the training step (models/bts/model/ray_sampler.py:108-160;
    configs/exp_kitti_raw.yaml: patch_size 8, ray_batch_size 2048 = 32 patches), different per frame.  Ray order: patch-major,
    row-major inside a patch (the layout ``PatchRaySampler.reconstruct`` views as (n, P, 8, 8))."""
    n = wl.rays.shape[0]
    g = torch.Generator().manual_seed(seed + 1234)
    all_rays = wl.rays.view(n, n_views, H, W, 8)
    sel = []
    for b in range(n):
        v = torch.randint(0, n_views, (patches,), generator=g).tolist()
        y = torch.randint(0, H - patch + 1, (patches,), generator=g).tolist()
        x = torch.randint(0, W - patch + 1, (patches,), generator=g).tolist()
        sel.append(torch.cat([all_rays[b, v[i], y[i]:y[i] + patch, x[i]:x[i] + patch].reshape(patch * patch, 8) for i in range(patches)]))
    wl.rays = torch.stack(sel).contiguous()
    wl.u = torch.rand((n * patches * patch * patch, wl.renderer_conf["n_coarse"]), generator=g)
    return wl


def cfg4_re10k_nvs(H=256, W=384, K=128, seed=0, lin_out_bias=0.0, ray_subsample=None) -> Workload:
    """BASELINE config 4: RE10K shape, C=32, Hd=32, one ResNet block, distance code, novel pose, 2 colour views."""
    poses = [_pose(), _pose(tx=0.15, yaw_deg=-2.0), _pose(tx=-0.1, tz=0.2, yaw_deg=1.5)]
    mc = kitti_model_conf(learn_empty=False, d_hidden=32, n_blocks=1, d_out=32, z_near=1.0, z_far=100.0,
                          code_mode="distance")
    rc = kitti_renderer_conf(K, hard_alpha_cap=False)
    novel = [_pose(tx=0.2, tz=-0.3, yaw_deg=3.0)]
    return make_workload(f"cfg4_{H}x{W}_K{K}", H=H, W=W, poses=poses, K3=RE10K_K, model_conf=mc, renderer_conf=rc,
                         ids_encoder=[0], ids_render=[1, 2], ray_poses=novel, seed=seed, lin_out_bias=lin_out_bias,
                         ray_subsample=ray_subsample)


def cfg5_stress(R=4096 * 4096, H=192, W=640, K=256, d_hidden=128, seed=0) -> Workload:
    """BASELINE config 5 (roofline sweep): random pixels x random poses within +-1 m / +-5 deg of the source."""
    g = torch.Generator().manual_seed(seed)
    mc = kitti_model_conf(learn_empty=False, d_hidden=d_hidden)
    rc = kitti_renderer_conf(K)
    wl = make_workload(f"cfg5_R{R}_K{K}_Hd{d_hidden}", H=H, W=W, poses=[_pose()], model_conf=mc, renderer_conf=rc,
                       ids_encoder=[0], ids_render=[0], ray_views=[0], ray_subsample=1, seed=seed)
    # random rays: origins within +-1 m, directions = source-frustum pixel directions rotated by <= 5 deg
    o = (torch.rand((R, 3), generator=g) * 2 - 1)
    px = torch.rand((R, 2), generator=g) * 2 - 1
    Kt = torch.tensor(KITTI_RAW_K)
    d = torch.stack(((px[:, 0] - Kt[0, 2]) / Kt[0, 0], (px[:, 1] - Kt[1, 2]) / Kt[1, 1], torch.ones(R)), -1)
    ang = (torch.rand((R, 2), generator=g) * 2 - 1) * math.radians(5.0)
    cy, sy, cp, sp = torch.cos(ang[:, 0]), torch.sin(ang[:, 0]), torch.cos(ang[:, 1]), torch.sin(ang[:, 1])
    dx = cy * d[:, 0] + sy * d[:, 2]
    dz = -sy * d[:, 0] + cy * d[:, 2]
    dy = cp * d[:, 1] - sp * dz
    dz = sp * d[:, 1] + cp * dz
    d = torch.stack((dx, dy, dz), -1)
    d = d / d.norm(dim=-1, keepdim=True)
    nf = torch.tensor([mc["z_near"], mc["z_far"]]).view(1, 2).expand(R, 2)
    wl.rays = torch.cat((o, d, nf), -1).view(1, R, 8).contiguous()
    wl.u = None  # too large to inject; perf mode uses the in-kernel generator
    return wl


# --------------------------------------------------------------------------------------------------------------------
# Model construction for synthetic workloads (tests, bench, smoke)
# --------------------------------------------------------------------------------------------------------------------

class FeatureEncoder(torch.nn.Module):
    """Encoder stand-in that returns a fixed feature tensor (the reference's own overfit harness does the same:
    ``models/bts/trainer_overfit.py:24-33``).  The render path only consumes the encoder's output layout
    (n*nv, C, H, W); the ResNet itself is out of scope."""

    def __init__(self, feats: torch.Tensor):
        super().__init__()
        self.feats = torch.nn.Parameter(feats.clone(), requires_grad=False)
        self.latent_size = feats.shape[1]
        self.scales = [0]

    def forward(self, x):
        return [self.feats]


def build_model(wl: Workload, device, state_dict=None):
    """(net, wrapped_renderer) for a workload; weights from ``state_dict`` (reference keys) or the reference's init
    scheme under ``wl.weight_seed``."""
    from . import BTSNet, NeRFRenderer
    torch.manual_seed(wl.weight_seed)
    net = BTSNet(wl.model_conf, encoder=FeatureEncoder(wl.feats))
    if state_dict is not None:
        missing, unexpected = net.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
        assert all(k.startswith("encoder.") for k in missing), missing
    else:
        with torch.no_grad():
            net.mlp_coarse.lin_out.bias.fill_(wl.lin_out_bias)
    renderer = NeRFRenderer.from_conf(wl.renderer_conf)
    net = net.to(device)
    wrapped = renderer.bind_parallel(net, gpus=None).eval().to(device)
    return net, wrapped
