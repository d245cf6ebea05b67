"""Discrete noise-level tables (reference: k_diffusion/external.py:41-84).

These are O(table) one-dimensional torch ops on whatever device the table lives on; they are
issued in the same order as the reference so `sigma_to_t` indices are bit-identical and the float
results match to the last bit on the same device.
"""
import torch
from torch import nn

from . import sampling


class DiscreteSchedule(nn.Module):
    """A mapping between continuous noise levels (sigmas) and a list of discrete noise levels."""

    def __init__(self, sigmas, quantize):
        super().__init__()
        self.register_buffer('sigmas', sigmas)
        self.register_buffer('log_sigmas', sigmas.log())
        self.quantize = quantize

    @property
    def sigma_min(self):
        return self.sigmas[0]

    @property
    def sigma_max(self):
        return self.sigmas[-1]

    def get_sigmas(self, n=None):
        if n is None:
            return sampling.append_zero(self.sigmas.flip(0))
        last = len(self.sigmas) - 1
        return sampling.append_zero(self.t_to_sigma(torch.linspace(last, 0, n, device=self.sigmas.device)))

    def sigma_to_t(self, sigma, quantize=None):
        if quantize is None:
            quantize = self.quantize
        log_sigma = sigma.log()
        dists = log_sigma - self.log_sigmas[:, None]             # [table, queries]
        if quantize:
            return dists.abs().argmin(dim=0).view(sigma.shape)   # nearest table entry (int64)
        lo_i = dists.ge(0).cumsum(dim=0).argmax(dim=0).clamp(max=self.log_sigmas.shape[0] - 2)
        hi_i = lo_i + 1
        lo, hi = self.log_sigmas[lo_i], self.log_sigmas[hi_i]
        w = ((lo - log_sigma) / (lo - hi)).clamp(0, 1)
        return ((1 - w) * lo_i + w * hi_i).view(sigma.shape)

    def t_to_sigma(self, t):
        t = t.float()
        lo_i, hi_i, w = t.floor().long(), t.ceil().long(), t.frac()
        return ((1 - w) * self.log_sigmas[lo_i] + w * self.log_sigmas[hi_i]).exp()
