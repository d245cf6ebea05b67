"""Backend switches (the reference's env flags live in models/flags.py:9-14).

K_DIFFUSION_USE_COMPILE / K_DIFFUSION_USE_FLASH_2 are accepted and ignored: there is no
torch.compile or flash-attn path here.  The only switch is the arithmetic of the token stream:

    KDB200_PRECISION = auto | fp32 | bf16      (default auto)

auto = bf16 when the call happens under torch.autocast(bfloat16) or the module's parameters are
bf16 (what `accelerate` mixed precision does for the reference), fp32 otherwise -- sample.py never
enables autocast, so the reference's inference default is true fp32 and so is ours.
"""
import os

import torch


def get_use_compile():
    return False


def get_use_flash_attention_2():
    return False


def resolve_precision(requested, param_dtype):
    req = (requested or os.environ.get("KDB200_PRECISION", "auto")).lower()
    if req in ("fp32", "float32"):
        return "fp32"
    if req in ("bf16", "bfloat16"):
        return "bf16"
    if req != "auto":
        raise ValueError(f"unknown precision {req!r}")
    if param_dtype == torch.bfloat16:
        return "bf16"
    if torch.is_autocast_enabled() and torch.get_autocast_gpu_dtype() == torch.bfloat16:
        return "bf16"
    return "fp32"
