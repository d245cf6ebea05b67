#!/usr/bin/env python
"""Golden outputs of the adaptive / fixed-step DPM-Solver entry points, recorded from the REAL reference
(build container only):

    python oracle/make_golden_dpm.py        # -> tests/golden/toy_dpm_solvers.npz

reference: k_diffusion/sampling.py:303-330 (PIDStepSizeController), :333-488 (DPMSolver), :491-516 (sample_dpm_fast / sample_dpm_adaptive).
Same toy denoiser and recorded-noise recipe as make_golden_next.py.  `log_likelihood` (:280-301) needs torchdiffeq, which is absent
from the reference tree and from this image: see make_golden_ll.py."""
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
    g = torch.Generator().manual_seed(12)
    x = torch.randn(3, 2, 5, 5, generator=g) * 80
    noise = [torch.randn(3, 2, 5, 5, generator=g) for _ in range(64)]

    def sampler():
        it = iter(noise)
        return lambda a, b: next(it)

    out, info = {}, {}
    for n in (4, 5, 6, 9, 10):            # nfe % 3 = 1, 2, 0: every order pattern of dpm_solver_fast (:412-417)
        out[f"dpm_fast_n{n}"] = S.sample_dpm_fast(toy2, x, 1e-2, 80., n, disable=True)
    out["dpm_fast_n7_eta05"] = S.sample_dpm_fast(toy2, x, 1e-2, 80., 7, disable=True, eta=0.5, s_noise=0.9, noise_sampler=sampler())
    out["dpm_fast_n6_eta1"] = S.sample_dpm_fast(toy2, x, 1e-2, 80., 6, disable=True, eta=1.0, noise_sampler=sampler())
    for name, kw in (("dpm_adaptive_o3", dict()), ("dpm_adaptive_o2", dict(order=2)), ("dpm_adaptive_o3_tight", dict(rtol=0.01, atol=0.002, h_init=0.1)),
                     ("dpm_adaptive_o3_pid", dict(pcoeff=0.2, icoeff=0.7, dcoeff=0.1, accept_safety=0.9)),
                     ("dpm_adaptive_o3_eta05", dict(eta=0.5, s_noise=0.9, noise_sampler=sampler()))):
        y, inf = S.sample_dpm_adaptive(toy2, x, 1e-2, 80., disable=True, return_info=True, **kw)
        out[name] = y
        info[name] = [inf["steps"], inf["nfe"], inf["n_accept"], inf["n_reject"]]
    np.savez(G.OUT / "toy_dpm_solvers.npz", x=x.numpy(), noise=torch.stack(noise).numpy(), **{k: v.numpy() for k, v in out.items()},
             **{k + "_info": np.array(v) for k, v in info.items()})
    print("wrote", G.OUT / "toy_dpm_solvers.npz", {k: tuple(v.shape) for k, v in out.items()}, info)


if __name__ == "__main__":
    main()
