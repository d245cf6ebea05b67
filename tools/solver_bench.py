#!/usr/bin/env python
"""Achieved HBM GB/s of the fused solver-step kernels (CUDA events, L2 flushed between iterations).

    python tools/solver_bench.py [--batch 32] [--json out.json]
Algorithmic bytes per fp32 element: euler 12 (read x, den; write x'), euler+noise 16, heun corrector 20, dpmpp_2m 20.
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "k-diffusion_b200"))
import torch

from k_diffusion import _native as N


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    shape = (a.batch, 3, 256, 256)
    x, d1, x2, d2, nz, out = (torch.randn(shape, device="cuda") for _ in range(6))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    n = x.numel()
    cases = {
        "euler_step (12 B/elem)": (12, lambda: N.euler_step(x, d1, -0.3, out=out)),
        "euler_step+noise (16 B/elem)": (16, lambda: N.euler_step(x, d1, -0.3, noise=nz, cn=0.5, out=out)),
        "heun_correct (20 B/elem)": (20, lambda: N.heun_correct(x, d1, x2, d2, -0.2, -0.4, out=out)),
        "dpmpp_2m_step (20 B/elem)": (20, lambda: N.dpmpp_2m_step(x, d1, d2, 0.7, -0.3, 1.4, -0.4, out=out)),
    }
    peaks = ROOT / "MEASURED_PEAKS.json"
    peak = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6650.0
    res = {}
    for name, (bpe, fn) in cases.items():
        for _ in range(3):
            fn()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        gbs = bpe * n / (ms * 1e-3) / 1e9
        res[name] = dict(us=ms * 1000, gbs=gbs, frac_of_peak=gbs / peak)
        print(f"{name:32s} {ms * 1000:8.1f} us  {gbs:7.0f} GB/s  ({gbs / peak:5.1%} of {'measured' if peaks.exists() else 'fallback'} {peak:.0f} GB/s)")
    if a.json:
        Path(a.json).write_text(json.dumps(dict(batch=a.batch, elements=n, peak_gbs=peak, results=res), indent=1))


if __name__ == "__main__":
    main()
