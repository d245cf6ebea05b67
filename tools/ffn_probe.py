#!/usr/bin/env python
"""What bounds the fused feed-forward kernel (tc_ffn_fused.cuh)?  Times it at the level-0 shape of the 256x256 model with
KDB200_FFN_DBG = 1 (weight chunks stream from L2 for the first tile only), 2 (no GEGLU arithmetic), 3 (both).  GPU box: python tools/ffn_probe.py"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "k-diffusion_b200"))
import torch

from k_diffusion import _native as N

lib = N.lib()
for M, F in [(131072, 384), (32768, 384)]:
    x = torch.randn(M, 128, device="cuda").to(torch.bfloat16)
    w_up = N.interleave_geglu_rows((torch.randn(2 * F, 128, device="cuda") / 128 ** 0.5).to(torch.bfloat16))
    w_dn = (torch.randn(128, F, device="cuda") / F ** 0.5).to(torch.bfloat16)
    ss = torch.zeros(M, 8, device="cuda")
    ss[:, 0] = x.float().pow(2).sum(1)
    for dbg in (0, 1, 2, 3):
        os.environ["KDB200_FFN_DBG"] = str(dbg)
        fn = lambda: N.check(lib.kdb_ffn_fused_bf16(N.ptr(x), N.ptr(w_up), N.ptr(w_dn), M, F, N.ptr(ss), None, N.stream()))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"M={M} d_ff={F} dbg={dbg}: {us:7.1f} us   {2.0 * M * 128 * 3 * F / us / 1e6:7.1f} TFLOP/s", flush=True)
os.environ["KDB200_FFN_DBG"] = "0"
