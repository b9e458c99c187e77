"""Drop-in ``BTSNet`` whose field evaluation runs in the sm_100a CUDA library (host mirror of
``models/bts/model/models_bts.py``).

Same constructor config keys, attributes, ``encode`` cache (``grid_f_*`` / ``grid_c_*``), ``set_scale`` and
``forward(xyz, coarse, viewdirs, far, only_density) -> (rgb, invalid, sigma)`` contract as the reference, and the same
state-dict keys (``encoder.*``, ``code_xyz._freqs/_phases``, ``mlp_coarse.*``, ``empty_feature``).  ``sample_features``
/ ``sample_colors`` + the MLP of the reference (``models_bts.py:138-338``) are replaced by
``bts_query_field`` / ``bts_render_fwd`` (``include/bts_b200.h``); the ResNet encoder stays a stock torch module.

There is NO torch/CPU implementation of the field here: without the CUDA library or on CPU tensors the calls raise.
Configurations the CUDA path does not implement (``sample_color=False``, ``combine_ids``, several encoder views,
``empty_empty``, ``use_viewdirs``) raise ``NotImplementedError`` instead of falling back (SURVEY.md §8b "Errors").
"""
import ctypes as C

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .code import PositionalEncoding
from .resnetfc import make_mlp

EPS = 1e-3


def _reference_backbone(conf):
    """The encoder is out of the hot path (BASELINE.json north_star: "the ResNet encoder remains a stock torch
    module"): when this class is dropped into the reference tree its own factory builds it."""
    try:
        from models.common.backbones.backbone_util import make_backbone  # the reference package, if on sys.path
    except Exception as e:  # pragma: no cover - depends on deployment
        raise RuntimeError("BTSNet needs an encoder: pass `encoder=` or run inside the reference tree so that "
                           "models.common.backbones.backbone_util.make_backbone is importable") from e
    return make_backbone(conf)


def _dev_f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.contiguous().float()
    return t


class _ScaleList:
    """``grid_f_features`` of the reference (a list with one (n, nv, C, H, W) map per encoder scale, models_bts.py:125-127)
    whose nearest-neighbour resize to the scale-0 size runs when a scale is first read, not for every scale on every encode."""

    def __init__(self, latents, shape):
        self._latents, self._shape, self._maps = list(latents), tuple(shape), {}

    def __len__(self):
        return len(self._latents)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        m = self._maps.get(i)
        if m is None:
            l = self._latents[i]
            n, nv, c, h, w = self._shape
            if tuple(l.shape[-2:]) != (h, w):
                l = F.interpolate(l, (h, w))
            m = self._maps[i] = l.view(n, nv, c, h, w)
        return m

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class BTSNet(nn.Module):
    def __init__(self, conf, encoder: nn.Module = None):
        super().__init__()
        self.d_min = conf.get("z_near")
        self.d_max = conf.get("z_far")
        self.learn_empty = conf.get("learn_empty", True)
        self.empty_empty = conf.get("empty_empty", False)
        self.inv_z = conf.get("inv_z", True)
        self.color_interpolation = conf.get("color_interpolation", "bilinear")
        self.code_mode = conf.get("code_mode", "z")
        if self.code_mode not in ("z", "distance"):
            raise NotImplementedError(f"Unknown mode for positional encoding: {self.code_mode}")
        self.flip_augmentation = conf.get("flip_augmentation", False)
        self.return_sample_depth = conf.get("return_sample_depth", False)
        self.sample_color = conf.get("sample_color", True)
        if not self.sample_color or self.empty_empty or self.return_sample_depth or self.color_interpolation != "bilinear":
            raise NotImplementedError("B200 render path implements sample_color=True, empty_empty=False, "
                                      "return_sample_depth=False, bilinear colours (all shipped configs)")

        self.encoder = encoder if encoder is not None else _reference_backbone(conf["encoder"])
        self.code_xyz = PositionalEncoding.from_conf(conf["code"], d_in=3)
        d_in = self.encoder.latent_size + self.code_xyz.d_out
        self._d_in, self._d_out = d_in, 1
        self.mlp_coarse = make_mlp(conf["mlp_coarse"], d_in, d_out=1)
        self.mlp_fine = make_mlp(conf["mlp_fine"], d_in, d_out=1, allow_empty=True)
        if self.learn_empty:
            self.empty_feature = nn.Parameter(torch.randn((self.encoder.latent_size,), requires_grad=True))
        self._scale = 0
        self._encode_gen = 0
        self._scene_cache = {}
        self._view_index_cache = {}
        self.grid_f_features = None

    # ------------------------------------------------------------------ reference API
    def set_scale(self, scale):
        self._scale = scale

    def get_scale(self):
        return self._scale

    def compute_grid_transforms(self, *args, **kwargs):
        pass

    def encode(self, images, Ks, poses_c2w, ids_encoder=None, ids_render=None, images_alt=None, combine_ids=None):
        """Boundary producer, kept in torch (models_bts.py:65-136): selects views, runs the encoder once, caches the
        feature maps / cameras / colour images the render kernels read."""
        if combine_ids is not None:
            raise NotImplementedError("combine_ids (view groups) is not implemented on the B200 render path")
        # torch.inverse (models_bts.py:66) = linalg.inv_ex + a host-side check of the LU status, i.e. a device sync per
        # frame that serialises the host with the previous step's kernels (2.8 of 4.6 ms of the training step's host time);
        # same factorisation, same result, no sync -- a singular camera pose yields inf/nan instead of an exception
        w2c = torch.linalg.inv_ex(poses_c2w, check_errors=False)[0]
        v = images.shape[1]
        ie = list(range(v)) if ids_encoder is None else list(ids_encoder)
        ir = list(range(v)) if ids_render is None else list(ids_render)
        take = self._take_views
        enc_in = take(images, ie)
        colour_src = images_alt if images_alt is not None else images * 0.5 + 0.5
        n, nv, c, h, w = enc_in.shape
        flip = bool(self.flip_augmentation and self.training and (torch.rand(1) > 0.5).item())
        x = enc_in.reshape(n * nv, c, h, w)
        if flip:
            x = torch.flip(x, dims=(-1,))
        latents = self.encoder(x)
        if flip:
            latents = [torch.flip(l, dims=(-1,)) for l in latents]
        h_, w_ = latents[0].shape[-2:]
        c_l = self.encoder.latent_size
        # models_bts.py:125-126 resizes EVERY latent to the scale-0 size on every encode; only the scale being rendered
        # is ever read (set_scale), so the resize happens on first access of a scale (scale 0 is already that size)
        self.grid_f_features = _ScaleList(latents, (n, nv, c_l, h_, w_))
        self.grid_f_Ks, self.grid_f_poses_w2c, self.grid_f_combine = take(Ks, ie), take(w2c, ie), None
        self.grid_c_imgs, self.grid_c_Ks, self.grid_c_poses_w2c, self.grid_c_combine = take(colour_src, ir), take(Ks, ir), take(w2c, ir), None
        # colour views that are the encoder view itself (eval_depth: ids_render == ids_encoder): the kernels reuse the
        # feature-view projection for them
        self._same_view_mask = 0
        if len(ie) == 1:
            for vi, r in enumerate(ir):
                if r == ie[0] and vi < 31:
                    self._same_view_mask |= 1 << vi
        self._encode_gen += 1
        self._scene_cache.clear()

    def _take_views(self, t, ids):
        """``t[:, ids]`` (models_bts.py:75-81, :128-135) without the host sync: indexing a CUDA tensor with a Python list
        uploads the index from pageable memory, which waits for every kernel already queued (3 ms of a 4.6 ms training
        step went there).  Consecutive ids become a slice copy, anything else an ``index_select`` with a cached device index."""
        ids = [int(i) for i in ids]
        if len(ids) > 0 and ids == list(range(ids[0], ids[0] + len(ids))):
            return t[:, ids[0]:ids[0] + len(ids)].clone()      # a fresh tensor, like advanced indexing returns
        key = (str(t.device), tuple(ids))
        idx = self._view_index_cache.get(key)
        if idx is None:
            idx = torch.tensor(ids, dtype=torch.long, device=t.device)
            self._view_index_cache[key] = idx
        return t.index_select(1, idx)

    def forward(self, xyz, coarse=True, viewdirs=None, far=False, only_density=False):
        """Point query (models_bts.py:266-338): xyz (n,P,3) -> rgb (n,P,3*Vc), invalid (n,P,Vc) float, sigma (n,P,1);
        with ``only_density`` rgb is zeros and invalid keeps the reference's (n,Vf,P,1) shape (models_bts.py:283,337)."""
        if viewdirs is not None:
            raise NotImplementedError("use_viewdirs is not part of the BTS field")
        lib = _lib.load()
        scene, keep = self._scene_struct()
        mlp, keep2 = self._mlp_struct()
        cfg, keep3 = self._cfg_struct(None)
        xyz_d = _dev_f32(xyz)
        self._require_cuda(xyz_d)
        n, P, _ = xyz_d.shape
        Vc, Cc = scene.Vc, scene.Cc
        dev = xyz_d.device
        sigma = torch.empty((n, P, 1), device=dev, dtype=torch.float32)
        inv_f = torch.empty((n, P), device=dev, dtype=torch.float32)
        stream = torch.cuda.current_stream(dev).cuda_stream
        if only_density:
            with _lib.on_device(dev):
                rc = lib.bts_query_field(C.byref(scene), C.byref(mlp), C.byref(cfg), xyz_d.data_ptr(), n, P, sigma.data_ptr(),
                                         inv_f.data_ptr(), None, None, stream)
            _lib.check(rc, "bts_query_field")
            return torch.zeros((n, P, Vc * 3), device=dev), inv_f.view(n, 1, P, 1), sigma
        rgb = torch.empty((n, P, Vc * Cc), device=dev, dtype=torch.float32)
        invalid = torch.empty((n, P, Vc), device=dev, dtype=torch.float32)
        with _lib.on_device(dev):
            rc = lib.bts_query_field(C.byref(scene), C.byref(mlp), C.byref(cfg), xyz_d.data_ptr(), n, P, sigma.data_ptr(),
                                     inv_f.data_ptr(), rgb.data_ptr(), invalid.data_ptr(), stream)
        _lib.check(rc, "bts_query_field")
        return rgb, invalid, sigma

    # ------------------------------------------------------------------ CUDA plumbing
    def invalidate_scene(self):
        """Drop the render-ready scene (folded map, packed colours): the next call rebuilds it."""
        self._scene_cache.clear()

    @staticmethod
    def _require_cuda(t):
        if not t.is_cuda:
            raise _lib.BtsCudaError("the BTS render path runs on CUDA tensors only (no CPU fallback)")

    def _mlp_struct(self):
        m = self.mlp_coarse
        s = _lib.BtsMlp()
        s.C, s.P, s.Hd, s.n_blocks = self.encoder.latent_size, self.code_xyz.d_out, m.d_hidden, m.n_blocks
        if m.n_blocks > _lib.BTS_MAX_BLOCKS:
            raise NotImplementedError(f"n_blocks={m.n_blocks} > {_lib.BTS_MAX_BLOCKS}")
        keep = [_dev_f32(m.lin_in.weight.detach()), _dev_f32(m.lin_in.bias.detach()),
                _dev_f32(m.lin_out.weight.detach()), _dev_f32(m.lin_out.bias.detach())]
        self._require_cuda(keep[0])
        s.w_in, s.b_in, s.w_out, s.b_out = (t.data_ptr() for t in keep)
        for i, blk in enumerate(m.blocks):
            ts = [_dev_f32(blk.fc_0.weight.detach()), _dev_f32(blk.fc_0.bias.detach()),
                  _dev_f32(blk.fc_1.weight.detach()), _dev_f32(blk.fc_1.bias.detach())]
            s.fc0_w[i], s.fc0_b[i], s.fc1_w[i], s.fc1_b[i] = (t.data_ptr() for t in ts)
            keep += ts
        return s, keep

    def _cfg_struct(self, renderer):
        c = _lib.BtsRenderCfg()
        code = self.code_xyz
        c.inv_z, c.code_mode, c.learn_empty = int(self.inv_z), 0 if self.code_mode == "z" else 1, int(self.learn_empty)
        c.num_freqs, c.include_input, c.freq_factor = code.num_freqs, int(code.include_input), code.freq_factor
        c.z_near, c.z_far = float(self.d_min), float(self.d_max)
        # python-double scalar arithmetic of models_bts.py:160-162, rounded to fp32 once like torch does
        c.inv_z_far = 1 / self.d_max
        c.inv_z_range = 1 / self.d_min - 1 / self.d_max
        c.z_range = self.d_max - self.d_min
        keep = []
        if renderer is not None:
            K = int(renderer.n_coarse)
            c.n_coarse, c.lindisp = K, int(renderer.lindisp)
            c.hard_alpha_cap, c.white_bkgd = int(renderer.hard_alpha_cap), int(renderer.white_bkgd)
            c.step = 1.0 / K
            dev = self.mlp_coarse.lin_in.weight.device
            steps = renderer._z_steps(K, dev)
            c.z_steps = steps.data_ptr()
            c.precision = int(getattr(renderer, "precision", 0))
            c.band_mode = 0 if getattr(renderer, "band", True) else 1
            seed_t = getattr(renderer, "seed_tensor", None)      # device int64 word mixed into the jitter seed (CUDA graphs)
            if seed_t is not None:
                c.seed_dev = seed_t.data_ptr()
                keep.append(seed_t)
            keep.append(steps)
        return c, keep

    def _scene_struct(self, sparse=False):
        """Render-ready scene for the current encode()/scale/weights; rebuilt when any of them changed.
        ``sparse=True`` (training): the folded map is allocated but NOT computed here -- ``_sparse_prepare`` folds the
        64-pixel tiles each ray batch samples (bts_mark_tiles + bts_fold_features_sparse); the tile mask lives with the
        scene."""
        if self.grid_f_features is None:
            raise RuntimeError("call encode() first")
        lib = _lib.load()
        feats = self.grid_f_features[self._scale]
        n, nv, Cf, H, W = feats.shape
        if nv != 1:
            raise NotImplementedError("B200 render path implements one encoder view (ids_encoder=[0], all shipped configs)")
        self._require_cuda(feats)
        w_in = self.mlp_coarse.lin_in.weight
        ef = self.empty_feature if self.learn_empty else None
        b_p = self.mlp_coarse.lin_in.bias
        key = (self._encode_gen, self._scale, w_in._version, w_in.data_ptr(), b_p._version, feats._version, feats.data_ptr(),
               None if ef is None else (ef._version, ef.data_ptr()))
        slot = ("sparse" if sparse else "scene", self._scale)  # one slot per scale (multiscale rendering, trainer.py:220-242)
        hit = self._scene_cache.get(slot)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        dev = feats.device
        stream = torch.cuda.current_stream(dev).cuda_stream
        Hd = self.mlp_coarse.d_hidden
        f = _dev_f32(feats.detach()).view(n, Cf, H, W)
        w = _dev_f32(w_in.detach())
        b_in = _dev_f32(self.mlp_coarse.lin_in.bias.detach())
        G = torch.empty((n, H, W, Hd), device=dev, dtype=torch.float32)
        g_empty = torch.empty((Hd,), device=dev, dtype=torch.float32) if ef is not None else None
        efd = _dev_f32(ef.detach()) if ef is not None else None
        mask = None
        if sparse:
            mask = torch.zeros((n * ((H * W + 63) // 64),), device=dev, dtype=torch.uint8)   # 0: no ray batch has needed the tile yet
        else:
            with _lib.on_device(dev):
                rc = lib.bts_fold_features(f.data_ptr(), n, Cf, H, W, w.data_ptr(), w.shape[1], b_in.data_ptr(), Hd,
                                           None if efd is None else efd.data_ptr(), G.data_ptr(),
                                           None if g_empty is None else g_empty.data_ptr(), stream)
            _lib.check(rc, "bts_fold_features")
        # operand-block copy of G for the band kernel (single-frame scenes: evaluation / novel-view rendering)
        G16 = g16_flags = None
        if n == 1 and Hd == 64 and getattr(self, "band_copy", True) and not sparse:
            G16 = torch.empty((lib.bts_g16_bytes(n, H, W, Hd),), device=dev, dtype=torch.uint8)
            g16_flags = torch.empty((1,), device=dev, dtype=torch.int32)
            with _lib.on_device(dev):
                rc = lib.bts_pack_g16(G.data_ptr(), n, H, W, Hd, G16.data_ptr(), g16_flags.data_ptr(), stream)
            _lib.check(rc, "bts_pack_g16")
        imgs = _dev_f32(self.grid_c_imgs.detach())
        _, Vc, Cc, Hc, Wc = imgs.shape
        img4 = torch.empty((n, Vc, Hc, Wc, 4), device=dev, dtype=torch.float32)
        if Vc > 0:
            with _lib.on_device(dev):
                rc = lib.bts_pack_images(imgs.data_ptr(), n * Vc, Cc, Hc, Wc, img4.data_ptr(), stream)
            _lib.check(rc, "bts_pack_images")
        Kf = _dev_f32(self.grid_f_Ks[:, 0, :3, :3].detach())
        Mf = _dev_f32(self.grid_f_poses_w2c[:, 0].detach())
        Kc = _dev_f32(self.grid_c_Ks[:, :, :3, :3].detach())
        Mc = _dev_f32(self.grid_c_poses_w2c.detach())
        s = _lib.BtsScene()
        s.n, s.H, s.W, s.Hd = n, H, W, Hd
        s.G, s.g_empty = G.data_ptr(), (None if g_empty is None else g_empty.data_ptr())
        s.K_f, s.w2c_f = Kf.data_ptr(), Mf.data_ptr()
        s.Vc, s.Cc, s.Hc, s.Wc = Vc, Cc, Hc, Wc
        s.img4, s.K_c, s.w2c_c = img4.data_ptr(), Kc.data_ptr(), Mc.data_ptr()
        s.same_view_mask = int(getattr(self, "_same_view_mask", 0)) if (Hc, Wc) == (H, W) else 0
        s.G16 = None if G16 is None else G16.data_ptr()
        s.g16_flags = None if g16_flags is None else g16_flags.data_ptr()
        keep = [f, w, b_in, G, g_empty, efd, imgs, img4, Kf, Mf, Kc, Mc, G16, g16_flags, mask]
        if sparse:
            s._sparse = dict(mask=mask, f=f, w=w, b_in=b_in, efd=efd, G=G, g_empty=g_empty, dims=(n, Cf, H, W, Hd))
        self._scene_cache[slot] = (key, s, keep)
        return s, keep

    def _sparse_ok(self, feats):
        """Sparse fold (training) applies when the tensor-core fold kernels take the shape and alignment of this map."""
        if not getattr(self, "sparse_fold", True) or feats is None or not feats.is_cuda:
            return False
        n, nv, Cf, H, W = feats.shape
        ok = nv == 1 and feats.dtype == torch.float32 and feats.is_contiguous() and feats.data_ptr() % 16 == 0
        return bool(ok and _lib.load().bts_sparse_fold_supported(Cf, H, W, self.mlp_coarse.d_hidden))

    def _sparse_prepare(self, renderer, rays, u, seed):
        """Fold the tiles of the current scale's map that this ray batch samples (and that no earlier batch of this
        encode() has folded): bts_mark_tiles runs the renderers' own sampling / projection / tap code."""
        import ctypes as C
        lib = _lib.load()
        scene, keep = self._scene_struct(sparse=True)
        st = scene._sparse
        cfg, keep3 = self._cfg_struct(renderer)
        rays_d = _dev_f32(rays)
        self._require_cuda(rays_d)         # no CPU fallback
        n, R, _ = rays_d.shape
        dev = rays_d.device
        u_d = _dev_f32(u.to(dev)) if u is not None else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        nb, Cf, H, W, Hd = st["dims"]
        with _lib.on_device(dev):
            rc = lib.bts_mark_tiles(C.byref(scene), C.byref(cfg), rays_d.data_ptr(), n, R,
                                    None if u_d is None else u_d.data_ptr(), int(seed) & (2 ** 64 - 1), st["mask"].data_ptr(), stream)
            _lib.check(rc, "bts_mark_tiles")
            rc = lib.bts_fold_features_sparse(st["f"].data_ptr(), nb, Cf, H, W, st["w"].data_ptr(), st["w"].shape[1],
                                              st["b_in"].data_ptr(), Hd, None if st["efd"] is None else st["efd"].data_ptr(),
                                              st["G"].data_ptr(), None if st["g_empty"] is None else st["g_empty"].data_ptr(),
                                              st["mask"].data_ptr(), stream)
            _lib.check(rc, "bts_fold_features_sparse")
        return scene, keep

    def render_rays(self, renderer, rays, u=None, seed=0, want_weights=False, want_alphas=False, want_z_samps=False,
                    want_rgb_samps=False, **extras):
        """Fused sampling + field + compositing for ``NeRFRenderer.forward`` (coarse pass).
        rays (n,R,8) -> dict of fresh tensors with the reference's shapes (nerf.py:377-401).  When gradients are
        enabled and the MLP / empty feature / encoder features require them, the call is recorded for autograd and
        ``loss.backward()`` runs ``bts_render_bwd`` (the reference differentiates its op chain with torch autograd,
        ``utils/base_trainer.py:297``)."""
        wants = dict(want_weights=want_weights, want_alphas=want_alphas, want_z_samps=want_z_samps,
                     want_rgb_samps=want_rgb_samps)
        # the consumer-side output diet (no per-sample flags, invalid mass per ray) is also available when training: both
        # are non-differentiable in the reference too (loss.py:105-108 only thresholds them)
        diet = {k: extras[k] for k in ("want_invalid", "want_invalid_mass") if k in extras}
        if torch.is_grad_enabled() and len(diet) == len(extras):
            feats = self.grid_f_features[self._scale] if self.grid_f_features is not None else None
            m = self.mlp_coarse
            diff = [feats, m.lin_in.weight, m.lin_in.bias, m.lin_out.weight, m.lin_out.bias,
                    self.empty_feature if self.learn_empty else None]
            blk = list(m.blocks[0].fc_0.parameters()) + list(m.blocks[0].fc_1.parameters()) if m.n_blocks == 1 else []
            if any(t is not None and t.requires_grad for t in diff + blk):
                if not (m.n_blocks == 0 and m.d_hidden <= 128) and not (m.n_blocks == 1 and m.d_hidden == 32):
                    raise NotImplementedError("backward of the B200 render path implements n_blocks=0 (d_hidden<=128) and "
                                              "n_blocks=1 with d_hidden=32 (the shipped KITTI / RE10K training configs)")
                keys = ["rgb", "depth"] + [k for k, w in (("invalid", diet.get("want_invalid", True)),
                                                          ("invalid_mass", diet.get("want_invalid_mass", False)),
                                                          ("weights", want_weights), ("alphas", want_alphas),
                                                          ("z_samps", want_z_samps), ("rgb_samps", want_rgb_samps)) if w]
                wants = dict(wants, **diet)
                ef = diff[5] if diff[5] is not None else torch.zeros(0, device=rays.device)
                if m.n_blocks == 1:
                    b0 = m.blocks[0]
                    blk_t = (b0.fc_0.weight, b0.fc_0.bias, b0.fc_1.weight, b0.fc_1.bias)
                else:
                    z = torch.zeros(0, device=rays.device)
                    blk_t = (z, z, z, z)
                outs = _RenderFunction.apply(self, renderer, rays, u, seed, wants, tuple(keys), diff[0], diff[1], diff[2],
                                             diff[3], diff[4], ef, *blk_t)
                return dict(zip(keys, outs))
        if extras and torch.is_grad_enabled() and any(p.requires_grad for p in self.mlp_coarse.parameters()):
            raise NotImplementedError("in-kernel ray generation / depth_z are inference-only: call them under "
                                      "torch.no_grad()")
        return self._render_rays_raw(renderer, rays, u, seed, **wants, **extras)

    def _render_rays_raw(self, renderer, rays, u=None, seed=0, want_weights=False, want_alphas=False, want_z_samps=False,
                         want_rgb_samps=False, want_invalid=True, want_invalid_mass=False, want_depth_z=False, gen=None,
                         want_sigma=False, sparse=False):
        """``rays`` (n,R,8), or ``rays=None`` with ``gen = (poses_c2w (n,v,4,4), Ks (n,v,3,3), H, W, z_near, z_far,
        norm_dir)`` to generate one ray per pixel of every view inside the kernel (bts_render_views_fwd)."""
        lib = _lib.load()
        scene, keep = self._sparse_prepare(renderer, rays, u, seed) if sparse else self._scene_struct()
        mlp, keep2 = self._mlp_struct()
        cfg, keep3 = self._cfg_struct(renderer)
        if gen is None:
            rays_d = _dev_f32(rays)
            self._require_cuda(rays_d)
            n, R, _ = rays_d.shape
            dev = rays_d.device
            if want_depth_z:
                raise ValueError("depth_z needs in-kernel ray generation (render_views)")
        else:
            poses, Ks, gh, gw, g_near, g_far, norm_dir = gen
            poses_d, Ks_d = _dev_f32(poses), _dev_f32(Ks[..., :3, :3])
            self._require_cuda(poses_d)
            n, gv = poses_d.shape[:2]
            R = gv * gh * gw
            dev = poses_d.device
            rg = _lib.BtsRayGen()
            rg.poses_c2w, rg.K, rg.v, rg.H, rg.W = poses_d.data_ptr(), Ks_d.data_ptr(), gv, gh, gw
            rg.norm_dir, rg.z_near, rg.z_far = int(norm_dir), float(g_near), float(g_far)
        K, Vc, Cc = cfg.n_coarse, scene.Vc, scene.Cc
        new = lambda *shape: torch.empty(shape, device=dev, dtype=torch.float32)  # noqa: E731
        ret = {"rgb": new(n, R, Vc * Cc), "depth": new(n, R)}
        if want_invalid:
            ret["invalid"] = new(n, R, K, Vc)
        if want_invalid_mass:
            ret["invalid_mass"] = new(n, R, Vc)
        if want_depth_z:
            ret["depth_z"] = new(n, R)
        if want_sigma:
            ret["sigma"] = new(n, R, K)
        if want_weights:
            ret["weights"] = new(n, R, K)
        if want_alphas:
            ret["alphas"] = new(n, R, K)
        if want_z_samps:
            ret["z_samps"] = new(n, R, K)
        if want_rgb_samps:
            ret["rgb_samps"] = new(n, R, K, Vc * Cc)
        out = _lib.BtsRenderOut()
        for k, t in ret.items():
            setattr(out, k, t.data_ptr())
        if scene.G16:      # work list of the band kernel (groups it hands to the gather kernel); [0] = their count
            ints = _lib.band_scratch_ints(n * R)
            self._band_work = torch.empty((ints,), device=dev, dtype=torch.int32)
            out.scratch, out.scratch_ints = self._band_work.data_ptr(), ints
        u_d = None
        if u is not None:
            u_d = _dev_f32(u.to(dev))
            if u_d.numel() != n * R * K:
                raise ValueError(f"u must have n*R*K = {n * R * K} elements, got {u_d.numel()}")
        stream = torch.cuda.current_stream(dev).cuda_stream
        if gen is None:
            with _lib.on_device(dev):
                rc = lib.bts_render_fwd(C.byref(scene), C.byref(mlp), C.byref(cfg), rays_d.data_ptr(), n, R,
                                        None if u_d is None else u_d.data_ptr(), int(seed) & (2 ** 64 - 1), C.byref(out), stream)
            _lib.check(rc, "bts_render_fwd")
        else:
            with _lib.on_device(dev):
                rc = lib.bts_render_views_fwd(C.byref(scene), C.byref(mlp), C.byref(cfg), C.byref(rg), n,
                                              None if u_d is None else u_d.data_ptr(), int(seed) & (2 ** 64 - 1), C.byref(out), stream)
            _lib.check(rc, "bts_render_views_fwd")
        return ret


class _RenderFunction(torch.autograd.Function):
    """Autograd node of the fused render: forward = bts_fold_features + bts_render_fwd, backward = bts_render_bwd
    (recompute per ray, scatter into dG) + the dense tail of the fold as two GEMMs."""

    @staticmethod
    def forward(ctx, net, renderer, rays, u, seed, wants, keys, feats, w_in, b_in, w_out, b_out, empty, fc0_w, fc0_b, fc1_w, fc1_b):
        with torch.no_grad():
            # training folds only the tiles of the map this batch samples (sparse fold), when the shape allows
            sparse = net._sparse_ok(feats)
            out = net._render_rays_raw(renderer, rays, u=u, seed=seed, want_sigma=True, sparse=sparse, **wants)
            # the scene this forward rendered (folded map of the current scale, cameras, colour views): the backward must
            # scatter into / recompute against exactly this one, whatever set_scale() / encode() did in between
            # (multiscale training renders scales 0..3 before one loss.backward(), trainer.py:220-242)
            ctx.scene, ctx.scene_keep = net._scene_struct(sparse=sparse)
            ctx.tile_mask = ctx.scene._sparse["mask"] if sparse else None
        ctx.net, ctx.renderer, ctx.seed, ctx.keys = net, renderer, seed, keys
        ctx.sigma = out.pop("sigma")
        ctx.has_u = u is not None
        ctx.has_block = fc0_w.numel() > 0
        ctx.K = int(renderer.n_coarse)
        ctx.save_for_backward(rays, u if u is not None else rays.new_zeros(0), feats, w_in, empty)
        res = tuple(out[k] for k in keys)
        ctx.mark_non_differentiable(*[out[k] for k in keys if k in ("invalid", "invalid_mass", "z_samps", "rgb_samps")])
        return res

    @staticmethod
    def backward(ctx, *grads):
        import ctypes as C
        net, renderer = ctx.net, ctx.renderer
        rays, u, feats, w_in, empty = ctx.saved_tensors
        lib = _lib.load()
        if int(renderer.n_coarse) != ctx.K:
            raise RuntimeError("renderer.n_coarse changed between forward and backward")
        scene, keep = ctx.scene, ctx.scene_keep             # forward-time scene (tensors kept alive by ctx)
        mlp, keep2 = net._mlp_struct()
        cfg, keep3 = net._cfg_struct(renderer)
        rays_d = _dev_f32(rays)
        n, R, _ = rays_d.shape
        dev = rays_d.device
        Hd, P, Cf = mlp.Hd, mlp.P, mlp.C
        gmap = dict(zip(ctx.keys, grads))
        up = _lib.BtsRenderGrad()
        held = []
        for name in ("rgb", "depth", "weights", "alphas"):
            g = gmap.get(name)
            if g is not None:
                g = _dev_f32(g)
                held.append(g)
                setattr(up, name, g.data_ptr())
        stream = torch.cuda.current_stream(dev).cuda_stream
        if ctx.tile_mask is not None:      # only the tiles the forward sampled receive gradient: zero those rows, leave the rest
            dG = torch.empty((scene.n, scene.H, scene.W, Hd), device=dev, dtype=torch.float32)
            with _lib.on_device(dev):
                rc = lib.bts_zero_tiles(dG.data_ptr(), scene.n, scene.H, scene.W, Hd, ctx.tile_mask.data_ptr(), stream)
            _lib.check(rc, "bts_zero_tiles")
        else:
            dG = torch.zeros((scene.n, scene.H, scene.W, Hd), device=dev, dtype=torch.float32)
        # every small accumulator lives in one zero-filled buffer (one launch instead of nine)
        sizes = [Hd, Hd * P, Hd, 1, Hd * Cf, Hd] + ([Hd * Hd, Hd, Hd * Hd, Hd] if ctx.has_block else [])
        flat = torch.zeros((sum(sizes),), device=dev, dtype=torch.float32)
        parts = torch.split(flat, sizes)
        dge, dwpe, dwout, dbout, d_WF_acc, d_b_acc = parts[0], parts[1].view(Hd, P), parts[2], parts[3], parts[4].view(Hd, Cf), parts[5]
        pg = _lib.BtsParamGrad()
        pg.d_G, pg.d_g_empty, pg.d_w_pe, pg.d_w_out, pg.d_b_out = (t.data_ptr() for t in (dG, dge, dwpe, dwout, dbout))
        dblk = (None, None, None, None)
        if ctx.has_block:
            dblk = (parts[6].view(Hd, Hd), parts[7], parts[8].view(Hd, Hd), parts[9])
            pg.d_fc0_w, pg.d_fc0_b, pg.d_fc1_w, pg.d_fc1_b = (t.data_ptr() for t in dblk)
        u_d = _dev_f32(u.to(dev)) if ctx.has_u else None
        stream = torch.cuda.current_stream(dev).cuda_stream
        with _lib.on_device(dev):
            rc = lib.bts_render_bwd(C.byref(scene), C.byref(mlp), C.byref(cfg), rays_d.data_ptr(), n, R,
                                    None if u_d is None else u_d.data_ptr(), int(ctx.seed) & (2 ** 64 - 1),
                                    ctx.sigma.data_ptr(), C.byref(up), C.byref(pg), stream)
        _lib.check(rc, "bts_render_bwd")
        # dense tail of the fold G = F W_F^T + b_in: bts_fold_features_bwd
        nb, _, _, H, W = feats.shape
        W_F = w_in.detach()[:, :Cf]
        need_df = ctx.needs_input_grad[7]
        feats_c = feats.detach()
        w_c = w_in.detach()
        if feats_c.dtype != torch.float32 or not feats_c.is_contiguous():
            feats_c = feats_c.contiguous().float()
        if w_c.dtype != torch.float32 or w_c.stride(1) != 1:
            w_c = w_c.contiguous().float()
        d_feats = torch.empty_like(feats_c) if need_df else None
        d_WF, d_b = d_WF_acc, d_b_acc
        with _lib.on_device(dev):
            if ctx.tile_mask is not None:
                rc = lib.bts_fold_features_bwd_sparse(dG.data_ptr(), feats_c.data_ptr(), nb, Cf, H, W, w_c.data_ptr(), w_c.stride(0), Hd,
                                                      d_feats.data_ptr() if need_df else None, d_WF.data_ptr(), d_b.data_ptr(),
                                                      ctx.tile_mask.data_ptr(), stream)
            else:
                rc = lib.bts_fold_features_bwd(dG.data_ptr(), feats_c.data_ptr(), nb, Cf, H, W, w_c.data_ptr(), w_c.stride(0), Hd,
                                               d_feats.data_ptr() if need_df else None, d_WF.data_ptr(), d_b.data_ptr(), stream)
        _lib.check(rc, "bts_fold_features_bwd")        # tensor-core kernel, or the CUDA-core one for odd shapes: no torch fallback
        d_b_in = d_b + dge
        if d_feats is not None and d_feats.dtype != feats.dtype:
            d_feats = d_feats.to(feats.dtype)
        d_empty = None
        if empty.numel() > 0:
            d_WF = d_WF + torch.outer(dge, empty.detach())
            d_empty = W_F.t() @ dge
        d_w_in = torch.cat((d_WF, dwpe), dim=1)
        return (None, None, None, None, None, None, None, d_feats, d_w_in, d_b_in, dwout.view(1, Hd), dbout, d_empty) + dblk
