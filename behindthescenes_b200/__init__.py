"""behindthescenes_b200 -- B200-native render hot path of Brummi/BehindTheScenes behind the reference's own API.

Public surface (names, signatures and dict contracts follow the reference):
    BTSNet, NeRFRenderer, PositionalEncoding, ResnetFC, make_mlp
The CUDA library (``lib/libbts_b200.so``, C ABI in ``include/bts_b200.h``) is loaded lazily on first use and is
mandatory: there is no CPU / eager fallback.
"""
from .code import PositionalEncoding  # noqa: F401
from .resnetfc import ResnetFC, ResnetBlockFC, make_mlp  # noqa: F401
from .bts_net import BTSNet  # noqa: F401
from .renderer import NeRFRenderer, _RenderWrapper  # noqa: F401

__all__ = ["BTSNet", "NeRFRenderer", "PositionalEncoding", "ResnetFC", "ResnetBlockFC", "make_mlp"]
