"""ResnetFC parameter container (host mirror of ``models/common/model/resnetfc.py`` + ``mlp_util.py``).

Keeps the reference's parameter names/shapes/initialisation (``lin_in``, ``blocks.{i}.fc_0/fc_1``, ``lin_out``) so
state dicts are interchangeable and optimisers see ``nn.Parameter``s.  The render path reads the parameters
through raw device pointers at call time (``bts_net.BTSNet._mlp_struct``); ``forward`` exists for module-API parity.
Only what the shipped BTS configs use is supported (d_latent=0, ReLU, no SPADE); anything else raises.
"""
import torch
from torch import nn


def _kaiming_linear(i, o, zero_weight=False):
    lin = nn.Linear(i, o)
    nn.init.zeros_(lin.bias)
    if zero_weight:
        nn.init.zeros_(lin.weight)
    else:
        nn.init.kaiming_normal_(lin.weight, a=0, mode="fan_in")
    return lin


class ResnetBlockFC(nn.Module):
    def __init__(self, size):
        super().__init__()
        self.size_in = self.size_h = self.size_out = size
        self.fc_0 = _kaiming_linear(size, size)
        self.fc_1 = _kaiming_linear(size, size, zero_weight=True)   # resnetfc.py:40
        self.activation = nn.ReLU()
        self.shortcut = None

    def forward(self, x):
        return x + self.fc_1(self.activation(self.fc_0(self.activation(x))))


class ResnetFC(nn.Module):
    def __init__(self, d_in, d_out=4, n_blocks=5, d_latent=0, d_hidden=128, beta=0.0, combine_layer=1000,
                 combine_type="average", use_spade=False):
        super().__init__()
        if d_latent != 0 or beta > 0 or use_spade:
            raise NotImplementedError("B200 render path supports ResnetFC with d_latent=0, ReLU, no SPADE")
        self.d_in, self.d_out, self.n_blocks, self.d_latent, self.d_hidden = d_in, d_out, n_blocks, d_latent, d_hidden
        self.combine_layer, self.combine_type, self.use_spade = combine_layer, combine_type, use_spade
        self.lin_in = _kaiming_linear(d_in, d_hidden)
        self.lin_out = _kaiming_linear(d_hidden, d_out)
        self.blocks = nn.ModuleList([ResnetBlockFC(d_hidden) for _ in range(n_blocks)])
        self.activation = nn.ReLU()

    @classmethod
    def from_conf(cls, conf, d_in, **kwargs):
        return cls(d_in, n_blocks=conf.get("n_blocks", 5), d_hidden=conf.get("d_hidden", 128), beta=conf.get("beta", 0.0),
                   combine_layer=conf.get("combine_layer", 1000), combine_type=conf.get("combine_type", "average"),
                   use_spade=conf.get("use_spade", False), **kwargs)

    def forward(self, zx, combine_inner_dims=(1,), combine_index=None, dim_size=None):
        x = self.lin_in(zx)
        for blk in self.blocks:
            x = blk(x)
        return self.lin_out(self.activation(x))


def make_mlp(conf, d_in, d_latent=0, allow_empty=False, **kwargs):
    kind = conf.get("type", "mlp")
    if kind == "resnet":
        return ResnetFC.from_conf(conf, d_in, d_latent=d_latent, **kwargs)
    if kind == "empty" and allow_empty:
        return None
    raise NotImplementedError(f"Unsupported MLP type for the B200 render path: {kind}")
