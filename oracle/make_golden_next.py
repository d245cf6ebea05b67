#!/usr/bin/env python
"""Golden trajectories of the SURVEY 8(f) row-1 samplers, recorded from the REAL reference (build container only):

    python oracle/make_golden_next.py        # -> tests/golden/toy_next_samplers.npz

Same recipe as the toy fixtures of make_golden.py: a nonlinear, sigma-dependent toy denoiser exercises every coefficient
path; stochastic samplers draw from a recorded list of noise tensors (the reference's default BrownianTree needs torchsde,
which is absent, so a noise_sampler is always passed -- that part of the default stays "parity unpinned")."""
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
    S = K.sampling
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 2, 5, 5, generator=g) * 80
    sig = S.get_sigmas_karras(10, 1e-2, 80)
    noise = [torch.randn(3, 2, 5, 5, generator=g) for _ in range(24)]

    def sampler():
        it = iter(noise)
        return lambda a, b: next(it)

    out = {
        "sample_dpm_2": S.sample_dpm_2(toy2, x, sig, disable=True),
        "sample_lms": S.sample_lms(toy2, x, sig, disable=True),
        "sample_lms_order2": S.sample_lms(toy2, x, sig, disable=True, order=2),
        "sample_dpm_2_ancestral": S.sample_dpm_2_ancestral(toy2, x, sig, disable=True, noise_sampler=sampler()),
        "sample_dpm_2_ancestral_eta05": S.sample_dpm_2_ancestral(toy2, x, sig, disable=True, eta=0.5, s_noise=0.9, noise_sampler=sampler()),
        "sample_dpmpp_2s_ancestral": S.sample_dpmpp_2s_ancestral(toy2, x, sig, disable=True, noise_sampler=sampler()),
        "sample_dpmpp_2s_ancestral_eta0": S.sample_dpmpp_2s_ancestral(toy2, x, sig, disable=True, eta=0., noise_sampler=sampler()),
        "sample_dpmpp_sde": S.sample_dpmpp_sde(toy2, x, sig, disable=True, noise_sampler=sampler()),
        "sample_dpmpp_sde_r03": S.sample_dpmpp_sde(toy2, x, sig, disable=True, eta=0.7, s_noise=0.9, r=0.3, noise_sampler=sampler()),
        "sample_dpmpp_2m_sde": S.sample_dpmpp_2m_sde(toy2, x, sig, disable=True, noise_sampler=sampler()),
        "sample_dpmpp_2m_sde_heun": S.sample_dpmpp_2m_sde(toy2, x, sig, disable=True, eta=0.6, solver_type="heun", noise_sampler=sampler()),
        "sample_dpmpp_2m_sde_eta0": S.sample_dpmpp_2m_sde(toy2, x, sig, disable=True, eta=0., noise_sampler=sampler()),
        "sample_dpmpp_3m_sde": S.sample_dpmpp_3m_sde(toy2, x, sig, disable=True, noise_sampler=sampler()),
        "sample_dpmpp_3m_sde_eta05": S.sample_dpmpp_3m_sde(toy2, x, sig, disable=True, eta=0.5, s_noise=0.8, noise_sampler=sampler()),
    }
    np.savez(G.OUT / "toy_next_samplers.npz", x=x.numpy(), sigmas=sig.numpy(), noise=torch.stack(noise).numpy(),
             **{k: v.numpy() for k, v in out.items()})
    print("wrote", G.OUT / "toy_next_samplers.npz", {k: tuple(v.shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
