#!/usr/bin/env python
"""KDB200_GEMM_TRACE=1 python tools/gemm_trace.py  -- per-tile role timestamps of the persistent GEMM (CTA 0)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "k-diffusion_b200"))
import torch
from k_diffusion import _native as N
for (M, Nn, K) in [(131072, 768, 128), (131072, 128, 128), (32768, 1536, 256)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = torch.randn(Nn, K, device="cuda").to(torch.bfloat16)
    for _ in range(2):
        c = N.gemm_bf16(a, w)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); c = N.gemm_bf16(a, w); torch.cuda.synchronize()
    print("M,N,K", M, Nn, K, "wall us (incl. trace sync)", (time.perf_counter() - t0) * 1e6, file=sys.stderr)
