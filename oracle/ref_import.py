"""Import the UNMODIFIED reference (Brummi/BehindTheScenes at /root/reference) as the parity anchor.

TEST INFRASTRUCTURE ONLY.  This module is used (a) by ``oracle/make_goldens.py`` to generate the committed
fixtures under ``tests/golden/`` and (b) by ``tests/test_oracle_vs_reference.py`` to pin the oracle port
(``oracle/bts_oracle.py``) against the real code.  ``/root/reference`` does not exist on the GPU box, so nothing
on the product path, in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import this file.

The reference needs three third-party modules that are not installed here (no network): ``dotmap``
(``models/common/render/nerf.py:9``), ``lpips`` (``models/bts/model/image_processor.py:5``) and ``omegaconf``
(``models/bts/model/ray_sampler.py:2``).  They are replaced by minimal stand-ins that only cover what the
hot path touches.  The ResNet encoder is replaced by a feature-tensor stand-in in the spirit of the
reference's own ``EncoderDummy`` (``models/bts/trainer_overfit.py:24-33``): the hot path only consumes its output.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("BTS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "bts", "model", "models_bts.py"))


class _AttrDict(dict):
    """dict with attribute access and ``toDict()`` -- the subset of dotmap.DotMap that nerf.py uses."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, _AttrDict):
                self[k] = _AttrDict(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def toDict(self):
        return {k: (v.toDict() if isinstance(v, _AttrDict) else v) for k, v in self.items()}


def _install_stubs():
    if "dotmap" not in sys.modules:
        m = types.ModuleType("dotmap")
        m.DotMap = _AttrDict
        sys.modules["dotmap"] = m
    if "lpips" not in sys.modules:
        m = types.ModuleType("lpips")

        class LPIPS(torch.nn.Module):  # never called on the hot path
            def __init__(self, *a, **kw):
                super().__init__()

        m.LPIPS = LPIPS
        sys.modules["lpips"] = m
    if "omegaconf" not in sys.modules:
        m = types.ModuleType("omegaconf")

        class ListConfig(list):
            pass

        m.ListConfig = ListConfig
        sys.modules["omegaconf"] = m


class FeatureStandIn(torch.nn.Module):
    """Encoder stand-in: returns a fixed (nv, C, H, W) feature tensor (one scale), like EncoderDummy."""

    def __init__(self, feats: torch.Tensor):
        super().__init__()
        self.feats = torch.nn.Parameter(feats.clone())
        self.latent_size = feats.shape[1]
        self.scales = [0]

    def forward(self, x):
        assert x.shape[0] == self.feats.shape[0], (x.shape, self.feats.shape)
        return [self.feats]


def load_reference():
    """Returns a namespace with the reference classes of the hot path."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.bts.model.models_bts as models_bts  # noqa
    import models.common.render.nerf as nerf  # noqa
    import models.bts.model.ray_sampler as ray_sampler  # noqa
    import models.common.util.util as util  # noqa
    ns = types.SimpleNamespace(models_bts=models_bts, nerf=nerf, ray_sampler=ray_sampler, util=util,
                               BTSNet=models_bts.BTSNet, NeRFRenderer=nerf.NeRFRenderer,
                               ImageRaySampler=ray_sampler.ImageRaySampler)
    return ns


def make_reference_net(ref, model_conf: dict, feats: torch.Tensor):
    """Build the reference BTSNet with the encoder replaced by ``FeatureStandIn(feats)``.

    ``feats``: (n*nv_enc, C, H, W).  The reference calls ``make_backbone(conf["encoder"])`` in its constructor
    (``models_bts.py:33``) which would build (and download) a ResNet; we swap the factory for the duration of
    the constructor only.
    """
    mb = ref.models_bts
    orig = mb.make_backbone
    mb.make_backbone = lambda conf, **kw: FeatureStandIn(feats)
    try:
        net = ref.BTSNet(model_conf)
    finally:
        mb.make_backbone = orig
    return net
