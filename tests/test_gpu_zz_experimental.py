"""SURVEY 8(f) row 1 on the GPU: the seven remaining fixed-schedule samplers.

Their oracle restatements are pinned against the reference and their host plans / executor are verified on the CPU
(tests/test_oracle_golden.py, tests/test_host_logic.py).  These GPU runs were written after the round's GPU budget was
spent, so they are marked xfail(strict=False): an XPASS in the round-end log is the first GPU evidence for these entry
points, a failure does not mask the north-star suite.  The marker goes away once they have been seen green.

The file name sorts after every other test module on purpose: should an experimental kernel ever fault, the sticky CUDA
error can only affect tests of this file.
"""
import pytest
import torch

import k_diffusion as K
from conftest import assert_close, load_npz
from oracle import kdiff_oracle as O
from test_gpu_parity import build

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180), pytest.mark.xfail(strict=False, reason="first GPU run of the 8f.1 entry points (no GPU minutes were left to confirm them)")]
S = K.sampling
DEV = "cuda"
toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)


def test_opaque_model_next_samplers_vs_reference_trajectories():
    z = load_npz("toy_next_samplers.npz")
    x, sig, nz = z["x"].to(DEV), z["sigmas"].to(DEV), z["noise"].to(DEV)

    def ns():
        it = iter(nz)
        return lambda a, b: next(it)

    runs = {
        "sample_dpm_2": lambda: S.sample_dpm_2(toy2, x, sig, disable=True),
        "sample_lms": lambda: S.sample_lms(toy2, x, sig, disable=True),
        "sample_lms_order2": lambda: S.sample_lms(toy2, x, sig, disable=True, order=2),
        "sample_dpm_2_ancestral": lambda: S.sample_dpm_2_ancestral(toy2, x, sig, disable=True, noise_sampler=ns()),
        "sample_dpmpp_2s_ancestral": lambda: S.sample_dpmpp_2s_ancestral(toy2, x, sig, disable=True, noise_sampler=ns()),
        "sample_dpmpp_sde_r03": lambda: S.sample_dpmpp_sde(toy2, x, sig, disable=True, eta=0.7, s_noise=0.9, r=0.3, noise_sampler=ns()),
        "sample_dpmpp_2m_sde": lambda: S.sample_dpmpp_2m_sde(toy2, x, sig, disable=True, noise_sampler=ns()),
        "sample_dpmpp_2m_sde_heun": lambda: S.sample_dpmpp_2m_sde(toy2, x, sig, disable=True, eta=0.6, solver_type="heun", noise_sampler=ns()),
        "sample_dpmpp_3m_sde_eta05": lambda: S.sample_dpmpp_3m_sde(toy2, x, sig, disable=True, eta=0.5, s_noise=0.8, noise_sampler=ns()),
    }
    for key, run in runs.items():
        assert_close(run(), z[key], rtol=1e-4, atol=2e-5, what=key)


def test_native_model_next_samplers_vs_oracle():
    """cfg1 (MNIST transformer, fp32 exact path): CUDA path against the CPU oracle on the same inputs, rtol 1e-3 / atol 1e-5."""
    cfg, sd, inner, model, z = build("cfg1_mnist")
    x, sigmas = z["x"].to(DEV), z["sigmas"].to(DEV)
    cc = z["class_cond"]
    ea = dict(class_cond=cc.to(DEV))
    oracle_model = O.make_denoiser(sd, cfg["model"])
    om = lambda xx, ss, **kw: oracle_model(xx, ss, class_cond=cc)
    assert_close(S.sample_dpm_2(model, x, sigmas, extra_args=ea, disable=True), O.sample_dpm_2(om, z["x"], z["sigmas"]), what="dpm_2")
    assert_close(S.sample_lms(model, x, sigmas, extra_args=ea, disable=True), O.sample_lms(om, z["x"], z["sigmas"]), what="lms")
    noise = z["noise"]
    it_g, it_o = iter(noise.to(DEV)), iter(noise)
    got = S.sample_dpmpp_2m_sde(model, x, sigmas, extra_args=ea, disable=True, noise_sampler=lambda a, b: next(it_g))
    want = O.sample_dpmpp_2m_sde(om, z["x"], z["sigmas"], lambda a, b: next(it_o))
    assert_close(got, want, what="dpmpp_2m_sde")


def test_brownian_default_graph_equals_eager(monkeypatch):
    """Unconditional native model + the default Brownian-tree noise: the captured graph must replay the eager launches."""
    cfg, sd, inner, model, z = build("sw64")
    x, sigmas = z["x"].to(DEV), z["sigmas"].to(DEV)
    ns = S.BrownianTreeNoiseSampler(x, float(sigmas[sigmas > 0].min()), float(sigmas.max()), seed=[3, 4][: x.shape[0]] if x.shape[0] <= 2 else list(range(x.shape[0])))
    S.clear_graph_cache()
    g1 = S.sample_dpmpp_2m_sde(model, x, sigmas, disable=True, noise_sampler=ns)
    monkeypatch.setenv("KDB200_CUDA_GRAPH", "0")
    e1 = S.sample_dpmpp_2m_sde(model, x, sigmas, disable=True, noise_sampler=ns)
    assert torch.isfinite(g1).all() and torch.equal(g1, e1)


@pytest.mark.parametrize("B,h,w,nh,shift", [(2, 16, 16, 2, 0), (2, 16, 16, 2, 4), (1, 8, 8, 4, 4), (3, 64, 64, 2, 4), (32, 64, 64, 2, 0), (8, 32, 32, 4, 4)])
def test_experimental_persistent_window_attention_matches_one_shot_kernel(monkeypatch, B, h, w, nh, shift):
    """KDB200_ATTN_PERSIST=1 (default off): double-buffered persistent variant of attn_tc_kernel<WINDOW>; same arithmetic, so the
    output must be bit-identical to the one-shot kernel's.  Its barrier waits are time-bounded (trap, not hang)."""
    from k_diffusion import _native as N_
    g = torch.Generator(device=DEV).manual_seed(B * h + w + nh + shift)
    qkv = (torch.randn(B, h * w, 3 * nh * 64, device=DEV, generator=g) * 0.2).to(torch.bfloat16)
    want = N_.attention(qkv, h, w, nh, 64, "shifted-window", 8, shift, fast=True)
    monkeypatch.setenv("KDB200_ATTN_PERSIST", "1")
    got = N_.attention(qkv, h, w, nh, 64, "shifted-window", 8, shift, fast=True)
    torch.cuda.synchronize()
    assert torch.equal(got, want)
