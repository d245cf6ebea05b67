#!/usr/bin/env python
"""Golden values of `log_likelihood`, recorded from the REAL reference (build container only):

    python oracle/make_golden_ll.py        # -> tests/golden/toy_log_likelihood.npz

reference: k_diffusion/sampling.py:280-301.  It integrates with `torchdiffeq.odeint(..., method='dopri5')`; torchdiffeq is neither in the
reference tree nor in this image, so the reference function is run with the ORACLE's dopri5 (`kdiff_oracle.odeint_dopri5`) installed as
`sampling.odeint`.  What this pins: the ODE right-hand side (autograd Hutchinson estimate), the state layout, the prior term and the
assembly of the result are the reference's own code.  What it cannot pin: torchdiffeq's accept / reject sequence (the integrator itself is
checked against a closed form and against scipy's RK45 in tests/test_oracle_golden.py)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

import make_golden as G
from oracle import kdiff_oracle as O


def odeint_like_torchdiffeq(func, y0, t, atol, rtol, method):
    """torchdiffeq's calling convention on top of the oracle's integrator: returns, per state member, the stack over `t`"""
    assert method == "dopri5" and len(t) == 2
    f = lambda s, y: func(torch.tensor(s, dtype=y0[0].dtype), y)          # torchdiffeq hands the function the time in the state's dtype
    y1, _ = O.odeint_dopri5(f, tuple(y0), float(t[0]), float(t[1]), atol, rtol)
    return tuple(torch.stack([a, b]) for a, b in zip(y0, y1))


def main():
    G._stub_missing()
    sys.path.insert(0, str(G.REF))
    import k_diffusion as K
    S = K.sampling
    S.odeint = odeint_like_torchdiffeq
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    gauss = lambda x, s, **kw: x * (0.49 / (0.49 + s[:, None, None, None] ** 2))          # exact denoiser of N(0, 0.7^2 I) data
    g = torch.Generator().manual_seed(21)
    x = torch.randn(3, 2, 6, 6, generator=g) * 0.7
    out = {"x": x}
    for name, model, kw in (("toy", toy2, {}), ("gauss", gauss, {}), ("toy_tight", toy2, dict(atol=1e-6, rtol=1e-6))):
        torch.manual_seed(77)
        v = torch.randint_like(x, 2) * 2 - 1          # what the reference will draw (:284) from the same generator state
        torch.manual_seed(77)
        ll, info = S.log_likelihood(model, x, 1e-2, 80., **kw)
        out[name + "_v"], out[name + "_ll"], out[name + "_fevals"] = v, ll, torch.tensor(info["fevals"])
        print(name, ll.tolist(), info)
    np.savez(G.OUT / "toy_log_likelihood.npz", **{k: t.numpy() for k, t in out.items()})
    print("wrote", G.OUT / "toy_log_likelihood.npz")


if __name__ == "__main__":
    main()
