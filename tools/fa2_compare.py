#!/usr/bin/env python
"""The library kernel the reference would call for global attention (flash_attn_qkvpacked_func, image_transformer_v2.py:383) against this
repo's pipelined kernel on the BASELINE global-attention shapes.  GPU box:  python tools/fa2_compare.py
flash-attn is LIBRARY code (not part of the product); it is timed here only as the kernel to beat."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "k-diffusion_b200"))
import torch

from k_diffusion import _native as N


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for name, B, h, w, nh in [("cfg2 mid level: S=256, 8 heads, batch 32", 32, 16, 16, 8), ("cfg5 mid level: S=1024, 16 heads, batch 16", 16, 32, 32, 16)]:
    S = h * w
    g = torch.Generator(device="cuda").manual_seed(0)
    t = torch.randn(B, S, 3, nh, 64, device="cuda", generator=g)
    t[:, :, :2] = t[:, :, :2] / t[:, :, :2].norm(dim=-1, keepdim=True) * 10 ** 0.5      # cosine-normalised q, k with scale 10
    qkv = t.to(torch.bfloat16)
    flat = qkv.reshape(B, S, 3 * nh * 64).contiguous()
    bound = torch.full([nh], 10.0, device="cuda")
    flops = 4.0 * S * S * 64 * nh * B
    ours = timed(lambda: N.attention(flat, h, w, nh, 64, "global", 0, 0, fast=True, logit_bound=bound))
    line = f"{name}: libkdb200 attn_pipe_kernel<GLOBAL> {ours:7.1f} us = {flops / ours / 1e6:6.1f} TFLOP/s"
    try:
        from flash_attn import flash_attn_qkvpacked_func
        fa = timed(lambda: flash_attn_qkvpacked_func(qkv, softmax_scale=1.0))
        ref = flash_attn_qkvpacked_func(qkv, softmax_scale=1.0).reshape(B, S, nh * 64).float()
        got = N.attention(flat, h, w, nh, 64, "global", 0, 0, fast=True, logit_bound=bound).float()
        line += f" | flash_attn_qkvpacked_func (flash-attn 2.8 library kernel) {fa:7.1f} us = {flops / fa / 1e6:6.1f} TFLOP/s | max |diff| {float((got - ref).abs().max()):.3e}"
    except Exception as exc:          # the library may have no kernel for this device
        line += f" | flash_attn unavailable: {exc!r}"
    print(line, flush=True)
