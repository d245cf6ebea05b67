#!/usr/bin/env python
"""Karras "churn" (s_churn > 0) of sample_euler / sample_heun / sample_dpm_2, recorded from the REAL reference (build container only):

    python oracle/make_golden_churn.py        # -> tests/golden/toy_churn.npz

reference: k_diffusion/sampling.py:117-135, :158-184, :187-214 (gamma, sigma_hat, eps * sqrt(sigma_hat^2 - sigma^2)).  The reference
draws `randn_like` on EVERY step, also when gamma = 0 (quirk Q1); the draws are captured per step so that an implementation which only
draws when gamma > 0 can be fed the same numbers."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np
import torch

import make_golden as G

CASES = {
    "euler_churn20": ("sample_euler", dict(s_churn=20.)),                                          # gamma capped at sqrt(2) - 1 on every step
    "heun_churn3_window": ("sample_heun", dict(s_churn=3., s_tmin=0.1, s_tmax=30., s_noise=1.1)),  # gamma = 0.3 inside the window only
    "dpm_2_churn2": ("sample_dpm_2", dict(s_churn=2.)),
}


def main():
    G._stub_missing()
    sys.path.insert(0, str(G.REF))
    import k_diffusion as K
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 2, 5, 5, generator=g) * 80
    sigmas = K.sampling.get_sigmas_karras(10, 1e-2, 80.)
    out = {"x": x, "sigmas": sigmas}
    real = torch.randn_like
    for name, (fn, kw) in CASES.items():
        draws = []

        def spy(t, *a, **k):
            draws.append(real(t, *a, **k))
            return draws[-1]

        torch.manual_seed(5)
        torch.randn_like = spy
        try:
            out[name] = getattr(K.sampling, fn)(toy2, x, sigmas, disable=True, **kw)
        finally:
            torch.randn_like = real
        out[name + "_eps"] = torch.stack(draws)
        print(name, len(draws), float(out[name].abs().max()))
    np.savez(G.OUT / "toy_churn.npz", **{k: v.numpy() for k, v in out.items()})
    print("wrote", G.OUT / "toy_churn.npz")


if __name__ == "__main__":
    main()
