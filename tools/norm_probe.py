#!/usr/bin/env python
"""Determinism probe for the fused-RMSNorm GEMM path: one model evaluation through sample_euler([sigma, 0]) (x_out == denoised),
run twice with the same input and once on a sub-batch; prints max |diff|."""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "k-diffusion_b200")):
    sys.path.insert(0, p)
import torch
import k_diffusion as K
os.environ["KDB200_CUDA_GRAPH"] = "0"
cfg = K.config.load_config(json.loads((ROOT / "tests/golden/cfg2_sw256_shapes.json").read_text())["config"])
inner = K.synth.synth_init_(K.config.make_model(cfg), seed=1).cuda().eval().set_precision("bf16")
model = K.Denoiser(inner, sigma_data=0.5)
x = K.parallel.init_noise(K.parallel.sample_seeds(3, 0, 32), (3, 256, 256), 4.0, "cuda")
sig = torch.tensor([3.0, 0.0], device="cuda")
run = lambda xx: K.sampling.sample_euler(model, xx, sig, disable=True)
a, b = run(x), run(x)
c = run(x[8:16].contiguous())
d = run(x[:1].contiguous())
print("fused_norm =", os.environ.get("KDB200_NO_FUSED_NORM", "0") != "1")
print("run-to-run max diff      :", float((a - b).abs().max()))
print("batch32 vs shard8 max diff:", float((a[8:16] - c).abs().max()), " rel", float((a[8:16] - c).norm() / c.norm()))
print("batch32 vs single max diff:", float((a[:1] - d).abs().max()))
print("finite:", bool(torch.isfinite(a).all()), "mean |out|:", float(a.abs().mean()))
