#!/usr/bin/env python
"""Which rows of layer0.qkv (v third) differ run-to-run, and by what per-row factor?"""
import json, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "k-diffusion_b200")):
    sys.path.insert(0, p)
import torch
import k_diffusion as K
cfg = K.config.load_config(json.loads((ROOT / "tests/golden/cfg2_sw256_shapes.json").read_text())["config"])
inner = K.synth.synth_init_(K.config.make_model(cfg), seed=1).cuda().eval().set_precision("bf16")
eng = inner.engine()
B = 32
x = torch.randn(B, 3, 256, 256, device="cuda") * 3
sig = torch.full([B], 3.0, device="cuda")
cond = eng.conditioning(sig[:1])
def tap(name):
    buf = eng.arm_tap(name, 131072 * 512, x.device)
    eng.forward(x, sig, cond, 0, 0.5, inner.resolved_precision())
    torch.cuda.synchronize()
    return buf[:eng.tap_count()].clone()
xin = tap("patch_in").view(-1, 128)
rstd_true = torch.rsqrt(xin.pow(2).mean(1) + 1e-6)
runs = [tap("layer0.qkv").view(-1, 384)[:, 256:] for _ in range(5)]
med = torch.stack(runs).median(0).values
for r, o in enumerate(runs):
    bad_rows = ((o - med).abs().amax(1) > 0).nonzero().flatten()
    print(f"run {r}: {bad_rows.numel()} rows differ from the median")
    for m in bad_rows[:10].tolist():
        num, den = o[m], med[m]
        k = den.abs().argmax()
        ratio = float(num[k] / den[k])
        want = ratio * float(rstd_true[m])
        tile, rit = m // 128, m % 128
        # does the wrong factor match the rstd of the same row in another m-tile?
        cands = rstd_true[rit::128]
        j = int((cands - want).abs().argmin())
        print(f"   row {m} (m-tile {tile}, row-in-tile {rit}): factor {ratio:.4f}; true rstd {float(rstd_true[m]):.4f}; implied rstd {want:.4f}; "
              f"closest same-row-in-tile rstd: m-tile {j} ({float(cands[j]):.4f}); delta tiles {j - tile}")
