"""Drop-in ``NeRFRenderer`` / ``_RenderWrapper`` (host mirror of ``models/common/render/nerf.py``).

Same constructor, ``from_conf``, ``bind_parallel``, ``sched_step``, buffers (``iter_idx``, ``last_sched``) and the same
live attributes callers mutate between calls (``n_coarse``, ``lindisp``, ``using_fine`` ... re-read on every call:
``scripts/images/gen_img_custom.py:111-112``, ``models/bts/trainer.py:224-229``).  ``forward`` keeps the reference's
signature and returns the same nested dict (``{"coarse": {"rgb", "depth", "invalid", [weights, alphas, z_samps,
rgb_samps]}}``, nerf.py:377-401) but stratified sampling, the chunked model loop and alpha compositing
(nerf.py:103-123, :210-313) run as ONE fused CUDA launch (``bts_render_fwd``) -- no (B*K, C) tensor is materialised.

Extensions over the reference signature (both optional, keyword-only in practice):
  ``u``     (n*R, K) stratified jitter in [0,1) to inject instead of drawing it (parity tests);
            by default the kernel draws it from a counter-based generator seeded from torch's CPU generator.
Importance / fine sampling (``n_fine > 0``, ``sample_from_dist``) is not used by any shipped config and raises.
"""
import torch

from . import _lib


class _Out(dict):
    """Nested dict with attribute access and ``toDict()`` -- what callers get from the reference's DotMap."""
    __getattr__ = dict.__getitem__

    def toDict(self):
        return {k: (v.toDict() if isinstance(v, _Out) else v) for k, v in self.items()}


class NeRFRenderer(torch.nn.Module):
    def __init__(self, n_coarse=128, n_fine=0, n_fine_depth=0, noise_std=0.0, depth_std=0.01, eval_batch_size=100000,
                 white_bkgd=False, lindisp=False, sched=None, hard_alpha_cap=False):
        super().__init__()
        self.n_coarse, self.n_fine, self.n_fine_depth = n_coarse, n_fine, n_fine_depth
        self.noise_std, self.depth_std = noise_std, depth_std
        self.eval_batch_size = eval_batch_size   # kept for config parity; the fused kernel needs no chunking
        self.white_bkgd, self.lindisp, self.hard_alpha_cap = white_bkgd, lindisp, hard_alpha_cap
        self.using_fine = n_fine > 0
        self.sched = sched if (sched is not None and len(sched) > 0) else None
        self.register_buffer("iter_idx", torch.tensor(0, dtype=torch.long), persistent=True)
        self.register_buffer("last_sched", torch.tensor(0, dtype=torch.long), persistent=True)
        self.precision = 0          # 0 = exact fp32 tier, 1 = fast tensor-core tier
        self._steps_cache = {}

    @classmethod
    def from_conf(cls, conf, white_bkgd=False, eval_batch_size=100000):
        return cls(conf.get("n_coarse", 128), conf.get("n_fine", 0), n_fine_depth=conf.get("n_fine_depth", 0),
                   noise_std=conf.get("noise_std", 0.0), depth_std=conf.get("depth_std", 0.01),
                   white_bkgd=conf.get("white_bkgd", white_bkgd), lindisp=conf.get("lindisp", True),
                   eval_batch_size=conf.get("eval_batch_size", eval_batch_size), sched=conf.get("sched", None),
                   hard_alpha_cap=conf.get("hard_alpha_cap", False))

    def _z_steps(self, K, device):
        """torch.linspace(0, 1 - 1/K, K) exactly as nerf.py:112-114 builds it (computed on the CPU once per K)."""
        key = (K, str(device))
        t = self._steps_cache.get(key)
        if t is None:
            t = torch.linspace(0, 1 - 1.0 / K, K).to(device)
            self._steps_cache[key] = t
        return t

    def sched_step(self, steps=1):
        if self.sched is None:
            return
        self.iter_idx += steps
        while self.last_sched.item() < len(self.sched[0]) and self.iter_idx.item() >= self.sched[0][self.last_sched.item()]:
            self.n_coarse = self.sched[1][self.last_sched.item()]
            self.n_fine = self.sched[2][self.last_sched.item()]
            self.last_sched += 1

    def forward(self, model, rays, want_weights=False, want_alphas=False, want_z_samps=False, want_rgb_samps=False,
                sample_from_dist=None, u=None, **extras):
        """``extras`` (inference-only extensions): ``want_invalid=False`` drops the (n,R,K,Vc) flag
        tensor, ``want_invalid_mass=True`` adds ``invalid_mass (n,R,Vc) = sum_k invalid*weights`` (the quantity
        ``loss.py:105-108`` / ``inference_setup.py:192`` threshold), ``gen=(poses_c2w, Ks, H, W, z_near, z_far,
        norm_dir)`` with ``rays=None`` generates one ray per pixel in the kernel, ``want_depth_z=True`` adds z-depth."""
        with torch.autograd.profiler.record_function("renderer_forward"):
            if self.sched is not None and self.last_sched.item() > 0:
                self.n_coarse = self.sched[1][self.last_sched.item() - 1]
                self.n_fine = self.sched[2][self.last_sched.item() - 1]
            assert rays is None or len(rays.shape) == 3
            if self.using_fine or self.n_fine > 0 or sample_from_dist is not None:
                raise NotImplementedError("fine / importance sampling is not implemented on the B200 render path "
                                          "(every shipped BTS config has n_fine = 0)")
            if self.training and self.noise_std > 0.0:
                raise NotImplementedError("sigma noise (noise_std > 0) is not implemented")
            if not hasattr(model, "render_rays"):
                raise TypeError("NeRFRenderer (B200) renders behindthescenes_b200.BTSNet fields only; there is no "
                                "generic per-point fallback loop")
            seed = int(torch.empty((), dtype=torch.int64).random_().item()) if u is None else 0
            coarse = model.render_rays(self, rays, u=u, seed=seed, want_weights=want_weights, want_alphas=want_alphas,
                                       want_z_samps=want_z_samps, want_rgb_samps=want_rgb_samps, **extras)
            return _Out(coarse=_Out(coarse))

    def bind_parallel(self, net, gpus=None, simple_output=False):
        if gpus is not None and len(gpus) > 1:
            raise NotImplementedError("single-process DataParallel is replaced by one process per GPU "
                                      "(behindthescenes_b200.rays.shard_rays + torch.distributed)")
        return _RenderWrapper(net, self, simple_output=simple_output)


class _RenderWrapper(torch.nn.Module):
    def __init__(self, net, renderer, simple_output):
        super().__init__()
        self.net, self.renderer, self.simple_output = net, renderer, simple_output

    def forward(self, rays, want_weights=False, want_alphas=False, want_z_samps=False, want_rgb_samps=False,
                sample_from_dist=None, u=None, **extras):
        if rays is not None and rays.shape[0] == 0:   # nerf.py:20-24
            return torch.zeros(0, 3, device=rays.device), torch.zeros(0, device=rays.device)
        so = self.simple_output
        out = self.renderer(self.net, rays, want_weights=want_weights and not so, want_alphas=want_alphas and not so,
                            want_z_samps=want_z_samps and not so, want_rgb_samps=want_rgb_samps and not so,
                            sample_from_dist=sample_from_dist, u=u, **extras)
        if so:
            return out.coarse.rgb, out.coarse.depth
        return out.toDict()

    def render_views(self, poses_c2w, Ks, height, width, z_near, z_far, norm_dir=True, **kwargs):
        """One ray per pixel of every given view, generated inside the kernel: replaces
        ``ImageRaySampler.sample`` + ``util.gen_rays`` + the ``(n, v*H*W, 8)`` ray tensor
        (``models/bts/model/ray_sampler.py:240-262``, ``models/common/util/util.py:244-273``); ``want_depth_z=True`` also
        returns the z-depth of ``utils/projection_operations.py:4-16``.  Output shapes are those of ``forward`` with
        R = v*H*W (ray index = (view*H + y)*W + x, the order ``ImageRaySampler.reconstruct`` expects)."""
        gen = (poses_c2w, Ks, int(height), int(width), float(z_near), float(z_far), bool(norm_dir))
        return self.forward(None, gen=gen, **kwargs)
