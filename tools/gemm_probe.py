#!/usr/bin/env python
"""What bounds the persistent tcgen05 GEMM?  Times the GEGLU and plain-store kernels at the cfg2 level shapes with parts of the epilogue
switched off (KDB200_GEMM_DBG, read per launch): 8 = the MMA issuers do not wait for the accumulator, 16 = nor for A / B, 32 = no TMA loads.
(Bits 1 / 2 / 4 = no staging stores / no epilogue math / no tcgen05.ld existed for profiles/r2_gemm_probe_epilogue_knockout.txt and were removed from the kernel.)
GPU box:  python tools/gemm_probe.py  [--trace]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "k-diffusion_b200"))
import torch

from k_diffusion import _native as N

lib = N.lib()
CLK = 1.965e9


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


MODES = [int(v) for v in os.environ.get("PROBE_MODES", "0,8,24,56").split(",")]
SHAPES = [(131072, 768, 128), (32768, 1536, 256), (8192, 3072, 512)] if os.environ.get("PROBE_ALL_SHAPES", "1") == "1" else [(131072, 768, 128), (32768, 1536, 256)]
print("KDB200_GEMM_MAX_NB =", os.environ.get("KDB200_GEMM_MAX_NB"), " modes: 8 MMA does not wait for the accumulator, "
      "16 MMA does not wait for A/B, 32 no TMA loads at all")
for (M, N2, K) in SHAPES:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N2, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    w_il = N.interleave_geglu_rows(w)
    out_g = torch.empty(M, N2 // 2, dtype=torch.bfloat16, device="cuda")
    out_s = torch.empty(M, N2, dtype=torch.bfloat16, device="cuda")
    tiles = (M // 128) * (N2 // 128)
    per_cta = tiles / 148
    for name, fn in (("geglu", lambda: N.check(lib.kdb_gemm_bf16_geglu(N.ptr(a), N.ptr(w_il), N.ptr(out_g), M, N2, K, None, N.stream()))),
                     ("store", lambda: N.check(lib.kdb_gemm_bf16(N.ptr(a), N.ptr(w), N.ptr(out_s), M, N2, K, N.stream())))):
        for dbg in MODES:
            os.environ["KDB200_GEMM_DBG"] = str(dbg)
            us = timed(fn)
            print(f"{name:6s} M={M:6d} N={N2:5d} K={K:4d} dbg={dbg}: {us:7.1f} us  {2.0 * M * N2 * K / us / 1e6:7.1f} TFLOP/s  "
                  f"{us * 1e-6 * CLK / per_cta:7.0f} cycles per 128x128 tile (MMA floor {K // 16 * 64})", flush=True)
    os.environ["KDB200_GEMM_DBG"] = "0"
