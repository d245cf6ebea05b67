#!/usr/bin/env python
"""`sample_lms` beyond the default order (reference sampling.py:247-277 accepts any order; its coefficients come from scipy quad),
recorded from the REAL reference (build container only):

    python oracle/make_golden_lms.py        # -> tests/golden/toy_lms_high_order.npz

Same toy denoiser, latent and schedule as make_golden_next.py (orders 4 and 2 live in toy_next_samplers.npz)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import numpy as np
import torch

import make_golden as G


def main():
    G._stub_missing()
    sys.path.insert(0, str(G.REF))
    import k_diffusion as K
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 2, 5, 5, generator=g) * 80
    sigmas = K.sampling.get_sigmas_karras(10, 0.05, 80.)
    out = {f"sample_lms_order{o}": K.sampling.sample_lms(toy2, x, sigmas, disable=True, order=o) for o in (5, 6, 7, 10)}
    np.savez(G.OUT / "toy_lms_high_order.npz", x=x.numpy(), sigmas=sigmas.numpy(), **{k: v.numpy() for k, v in out.items()})
    print("wrote", G.OUT / "toy_lms_high_order.npz", {k: float(v.abs().max()) for k, v in out.items()})


if __name__ == "__main__":
    main()
