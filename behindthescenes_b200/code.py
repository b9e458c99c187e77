"""Positional-code parameters (host mirror of ``models/common/model/code.py``).

On the render path the code is evaluated inside the CUDA kernels (``csrc/common.cuh::positional_code``); this module
only carries the configuration and the two buffers (``_freqs``, ``_phases``) so checkpoints keep the reference's
state-dict keys (``renderer.net.code_xyz._freqs/_phases``, shape (1, 2*num_freqs, 1)).
"""
import math

import torch


class PositionalEncoding(torch.nn.Module):
    def __init__(self, num_freqs=6, d_in=3, freq_factor=math.pi, include_input=True):
        super().__init__()
        self.num_freqs, self.d_in, self.include_input = int(num_freqs), int(d_in), bool(include_input)
        self.freq_factor = float(freq_factor)
        self.freqs = self.freq_factor * 2.0 ** torch.arange(0, self.num_freqs)
        self.d_out = 2 * self.num_freqs * self.d_in + (self.d_in if self.include_input else 0)
        self.register_buffer("_freqs", self.freqs.repeat_interleave(2).view(1, -1, 1))
        ph = torch.zeros(2 * self.num_freqs)
        ph[1::2] = math.pi * 0.5
        self.register_buffer("_phases", ph.view(1, -1, 1))

    @classmethod
    def from_conf(cls, conf, d_in=3):
        return cls(conf.get("num_freqs", 6), d_in, conf.get("freq_factor", math.pi), conf.get("include_input", True))

    def forward(self, x):
        """(B, d_in) -> (B, d_out).  Module-API parity only; the render kernels never call this."""
        e = torch.sin(torch.addcmul(self._phases, x.unsqueeze(1), self._freqs)).flatten(1)
        return torch.cat((x, e), -1) if self.include_input else e
