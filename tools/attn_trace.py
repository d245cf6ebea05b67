#!/usr/bin/env python
"""Role timelines of the pipelined attention kernel (CTA 0): KDB200_ATTN_TRACE=1 python tools/attn_trace.py  (GPU box).
Prints, per shape, the clock64 stamps of the producer / MMA issuer / two softmax groups (see tc_attention_pipe.cuh PA_TRACE)."""
import os
import sys
from pathlib import Path

os.environ.setdefault("KDB200_ATTN_TRACE", "1")
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "k-diffusion_b200"))
import torch

from k_diffusion import _native as N


def qkv(B, h, w, nh, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randn(B, h * w, 3, nh, 64, device="cuda", generator=g)
    t[:, :, :2] = t[:, :, :2] / t[:, :, :2].norm(dim=-1, keepdim=True) * 10 ** 0.5
    return t.to(torch.bfloat16).reshape(B, h * w, 3 * nh * 64)


for name, B, h, w, nh, kind, param, shift in [("global S=256 (cfg2 L2)", 32, 16, 16, 8, "global", 0, 0), ("global S=1024 (cfg5 L2)", 16, 32, 32, 16, "global", 0, 0),
                                              ("window L0 shifted (cfg2)", 32, 64, 64, 2, "shifted-window", 8, 4),
                                              ("window L0 unshifted (cfg2)", 32, 64, 64, 2, "shifted-window", 8, 0),
                                              ("neighbourhood L0 (cfg3 at B=32)", 32, 64, 64, 2, "neighborhood", 7, 0)]:
    x = qkv(B, h, w, nh)
    bound = torch.full([nh], 10.0, device="cuda")
    for _ in range(2):
        print(f"--- {name}", file=sys.stderr, flush=True)
        N.attention(x, h, w, nh, 64, kind, param, shift, fast=True, logit_bound=bound)
        torch.cuda.synchronize()
