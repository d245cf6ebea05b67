#!/usr/bin/env python
"""bf16 error budget, measured on the REAL reference (build container only):

    python oracle/make_golden_bf16.py        # -> tests/golden/bf16_budget.json

The north-star tolerance (rtol 1e-3 / atol 1e-5) is an fp32 gate.  BASELINE configs 2-5 run in bf16; the reference's own bf16
mode is `torch.autocast(bfloat16)` around the same module ("Reference precision modes").  This script runs the
reference module twice on identical inputs -- plain fp32 and under `torch.autocast('cpu', torch.bfloat16)` -- and records the
distance between the two: that number is the reference's OWN bf16 noise, and the GPU tests allow the CUDA bf16 path a stated
small multiple of it (tests/test_gpu_bf16_parity.py) instead of a hand-picked budget.

Recorded for: sw64 (forward B=2 + Heun 6 steps), cfg2 256x256 (forward B=1, several sigmas; Heun 10 Karras steps B=1), and the
cfg5 shape (512x512, widths 256/512/1024, shifted-window variant; Heun 2 steps B=1).
Latents and solver arithmetic stay fp32 in both runs (as in the reference's demo()/sample paths under accelerate).
"""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import torch

import make_golden as G


def dist(a, b):
    a, b = a.double(), b.double()
    return dict(rel_l2=float((a - b).norm() / b.norm()), max_abs=float((a - b).abs().max()), ref_rms=float(b.pow(2).mean().sqrt()))


def main():
    G._stub_missing()
    sys.path.insert(0, str(G.REF))
    import k_diffusion as K
    synth = G._load_synth()
    S = K.sampling
    torch.set_num_threads(8)

    def build(overrides=None):
        cfg = json.loads((G.REF / "configs" / "config_oxford_flowers_shifted_window.json").read_text())
        for k, v in (overrides or {}).items():
            cfg["model"][k] = v
        cfg = K.config.load_config(cfg)
        inner = K.config.make_model(cfg).eval().requires_grad_(False)
        base = inner.state_dict()
        inner.load_state_dict(synth.synth_state_dict({k: v.shape for k, v in base.items()}, seed=1, base=base))
        return K.config.make_denoiser_wrapper(cfg)(inner)

    def both(fn):
        with torch.no_grad():
            ref = fn()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                low = fn()
        return dist(low.float(), ref)

    out = {"how": "reference module fp32 vs the same module under torch.autocast('cpu', torch.bfloat16), identical inputs/weights "
                  "(synth seed 1); rel_l2 = |bf16 - fp32|_2 / |fp32|_2 over the whole output"}
    # sw64: the fixture inputs of tests/golden/sw64.npz (seed 124)
    model = build(dict(input_size=[64, 64]))
    g = torch.Generator().manual_seed(124)
    x = torch.randn(2, 3, 64, 64, generator=g) * 160
    sig_t = torch.tensor([0.3, 40.0])
    sigmas = S.get_sigmas_karras(6, 1e-2, 160)
    out["sw64_forward"] = both(lambda: model(x, sig_t))
    out["sw64_heun6"] = both(lambda: S.sample_heun(model, x, sigmas, disable=True))
    # cfg2 model, 256x256, B=1
    model = build()
    g = torch.Generator().manual_seed(125)
    x = torch.randn(1, 3, 256, 256, generator=g) * 160
    for s in (0.05, 2.5, 40.0):
        xs = x / 160 * (s * s + 0.25) ** 0.5           # a latent at noise level s (data std 0.5)
        out[f"cfg2_forward_sigma{s}"] = both(lambda: model(xs, torch.tensor([s])))
    sig10 = S.get_sigmas_karras(10, 1e-2, 160)
    out["cfg2_heun10"] = both(lambda: S.sample_heun(model, x, sig10, disable=True))
    # cfg5 shape (512x512, widths 256/512/1024), Heun 2 Karras steps, B=1: what bench.py's parity leg can afford on the CPU for cfg5.
    # The reference cannot run its neighbourhood layers here (natten is absent), so this is the shifted-window variant of the same
    # widths / depths / resolution: same GEMM shapes and accumulation lengths, i.e. the same bf16 noise sources.
    model = build(dict(input_size=[512, 512], widths=[256, 512, 1024], depths=[2, 2, 4]))
    g = torch.Generator().manual_seed(126)
    x = torch.randn(1, 3, 512, 512, generator=g) * 160
    sig2 = S.get_sigmas_karras(2, 1e-2, 160)
    out["cfg5shape_sw_heun2"] = both(lambda: S.sample_heun(model, x, sig2, disable=True))
    (G.OUT / "bf16_budget.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
