"""SURVEY 8(f) row 4, `log_likelihood` (reference sampling.py:280-301) on the GPU.

The oracle is pinned on the CPU (tests/test_oracle_golden.py: the reference function run with the oracle's dopri5, a closed form, scipy)
and the product's host loop is verified with stubbed kernels (tests/test_host_logic.py).  Here the CUDA path runs: the error-ratio
kernel against its formula, the autograd branch around opaque models against the values recorded from the reference and the Gaussian
closed form, and the native cfg1 model (finite-difference divergence on the engine's fp32 path) against the oracle's autograd evaluation.
"""
import math

import pytest
import torch

import k_diffusion as K
from conftest import load_npz
from oracle import kdiff_oracle as O
from test_gpu_parity import build

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
S = K.sampling
DEV = "cuda"
toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
gauss = lambda x, s, **kw: x * (0.49 / (0.49 + s[:, None, None, None] ** 2))


def test_rk_error_kernel_matches_formula():
    from k_diffusion import _native
    g = torch.Generator(device=DEV).manual_seed(9)
    y0 = torch.randn(4, 3, 64, 64, device=DEV, generator=g) * 3
    y1 = y0 + torch.randn(4, 3, 64, 64, device=DEV, generator=g) * 0.1
    err = torch.randn(4, 3, 64, 64, device=DEV, generator=g) * 1e-3
    want = float((err.double() / (1e-4 + 1e-4 * torch.maximum(y0.abs(), y1.abs()).double())).pow(2).mean().sqrt())
    got = _native.rk_error(err, y0, y1, 1e-4, 1e-4)
    assert abs(got - want) <= 1e-5 * want
    assert _native.rk_error(err, y0, y1, 1e-4, 1e-4) == got                    # deterministic reduction
    odd = slice(0, 4099)                                                        # a length that is no multiple of the vector width
    e1, a1, b1 = (t.flatten()[odd].contiguous() for t in (err, y0, y1))
    want = float((e1.double() / (1e-3 + 0.05 * torch.maximum(a1.abs(), b1.abs()).double())).pow(2).mean().sqrt())
    assert abs(_native.rk_error(e1, a1, b1, 1e-3, 0.05) - want) <= 1e-5 * want


@pytest.mark.parametrize("name,kw", [("toy", {}), ("gauss", {}), ("toy_tight", dict(atol=1e-6, rtol=1e-6))])
def test_opaque_model_log_likelihood_vs_reference_values(name, kw):
    z = load_npz("toy_log_likelihood.npz")
    model = gauss if name == "gauss" else toy2
    x = z["x"].to(DEV)
    ll, info = S.log_likelihood(model, x, 1e-2, 80., v=z[name + "_v"].to(DEV), **kw)
    want = z[name + "_ll"]
    assert ll.shape == (3,) and ll.is_cuda and info["fevals"] == 2 + 6 * (info["n_accept"] + info["n_reject"])
    # early error estimates sit at fp32 round-off, so a step more or less than the CPU run is legitimate; the value is not
    assert abs(info["fevals"] - int(z[name + "_fevals"])) <= 18, info
    tol = kw.get("rtol", 1e-4)
    assert float((ll.cpu() - want).abs().max()) <= 5 * tol * float(want.abs().max()), (ll, want)
    if name == "gauss":                                                       # closed form: sum log N(x_i; 0, s^2 + sigma_min^2)
        exact = torch.distributions.Normal(0, math.sqrt(0.49 + 1e-4)).log_prob(z["x"].double()).flatten(1).sum(1)
        assert float((ll.cpu().double() - exact).abs().max()) <= 1e-3 * float(exact.abs().max())


def test_log_likelihood_requires_a_differentiable_or_native_model():
    z = load_npz("toy_log_likelihood.npz")
    with pytest.raises(RuntimeError):
        S.log_likelihood(lambda x, s: x.detach() * 0.5, z["x"].to(DEV), 1e-2, 80.)
    with pytest.raises(RuntimeError):
        S.log_likelihood(toy2, z["x"], 1e-2, 80.)                               # CPU tensor: there is no CPU path


def test_native_model_log_likelihood_vs_oracle_autograd():
    """cfg1 (MNIST transformer, class-conditional): the engine has forward kernels only, so the Hutchinson quadratic form v^T J v is a
    4th-order central difference of fp32 engine evaluations; the oracle differentiates its own model with autograd."""
    cfg, sd, inner, model, z = build("cfg1_mnist")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 1, 28, 28, generator=g) * 0.4 + 0.1
    v = torch.randint(0, 2, x.shape, generator=g).float() * 2 - 1
    cc = torch.tensor([1, 9])
    ea = dict(class_cond=cc.to(DEV))
    om = O.make_denoiser(sd, cfg["model"])
    # the right-hand side at three noise levels
    rhs, count = S._likelihood_rhs(model, x.to(DEV), ea, v.to(DEV), 1e-2)
    for sigma in (0.02, 0.7, 30.0):
        xs = x * (1 + sigma)
        d, d_ll = rhs(sigma, (xs.to(DEV), torch.zeros(2, device=DEV)))
        with torch.enable_grad():
            xg = xs.clone().requires_grad_()
            dd = (xg - om(xg, torch.full((2,), sigma), class_cond=cc)) / sigma
            want = (v * torch.autograd.grad((dd * v).sum(), xg)[0]).flatten(1).sum(1)
        assert float((d.cpu() - dd.detach()).abs().max()) <= 1e-3 * float(dd.abs().max()) + 1e-4 * float(xs.abs().max()) / sigma
        # (measured with the oracle's fp32 model in place of the engine: |error of v^T J_D v| < 1e-3, i.e. 1e-3 / sigma here)
        assert float((d_ll.cpu() - want).abs().max()) <= 5e-2 / sigma + 1e-3 * float(want.abs().max()), (sigma, d_ll, want)
    assert count[0] == 3
    # the integral
    ll, info = S.log_likelihood(model, x.to(DEV), 1e-2, 80., extra_args=ea, v=v.to(DEV))
    ll_o, info_o = O.log_likelihood(om, x, 1e-2, 80., extra_args=dict(class_cond=cc), v=v)
    assert info["fevals"] == 2 + 6 * (info["n_accept"] + info["n_reject"]) and info["n_accept"] >= 10
    # at the default tolerances two correct integrations differ by a few rtol * |ll| (the tight-tolerance value lies between them)
    assert float((ll.cpu() - ll_o).abs().max()) <= 1.5e-3 * float(ll_o.abs().max()), (ll, ll_o, info, info_o)
