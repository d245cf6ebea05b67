"""Differential fuzzing of the host side: every fixed-schedule sampler of the product (step plans + executor, native primitives replaced
by torch one-liners -- test-only stubs) against the oracle (pinned to the reference) on 400 random cases: Karras / exponential /
polyexponential / hand-made schedules, with and without the trailing zero, 1 to 12 steps, three decades of sigma range, random
eta / s_noise / r / order / solver_type.  CPU only; fixed seed."""
import math
import random

import numpy as np
import pytest
import torch

import k_diffusion as K
from oracle import kdiff_oracle as O

S = K.sampling
f = np.float32
toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)


def _stub(monkeypatch):
    from k_diffusion import _native
    monkeypatch.setattr(_native, "require_cuda", lambda *t: None)
    monkeypatch.setattr(_native, "f32c", lambda t: t.to(torch.float32).contiguous())
    monkeypatch.setattr(_native, "lincomb", lambda ts, cs, out=None: sum(f(c) * t for t, c in zip(ts, cs)))
    monkeypatch.setattr(_native, "euler_step", lambda x, den, r, noise=None, cn=0.0, out=None: x + (x - den) * f(r) + (0 if noise is None else noise * f(cn)))
    monkeypatch.setattr(_native, "heun_correct", lambda x, d1, x2, d2, a1, a2, out=None: x + ((x - d1) * f(a1) + (x2 - d2) * f(a2)))
    monkeypatch.setattr(_native, "dpmpp_2m_step", lambda x, den, old, a, b, k1, k0, out=None:
                        f(a) * x - f(b) * (f(k1) * den + (f(k0) * old if old is not None else 0)))


def _schedule(rng):
    n = rng.randint(1, 12)
    kind = rng.choice(["karras", "exp", "poly", "manual", "nozero"])
    smin = 10 ** rng.uniform(-2.5, 0)
    smax = smin * 10 ** rng.uniform(0.3, 3.5)
    if kind == "karras":
        return O.get_sigmas_karras(n, smin, smax, rho=rng.uniform(1, 9))
    if kind == "exp":
        return O.get_sigmas_exponential(n, smin, smax)
    if kind == "poly":
        return O.get_sigmas_polyexponential(n, smin, smax, rho=rng.uniform(0.3, 2))

    def pick(m):          # irregular but separated nodes: nearly repeated sigmas make the multistep formulas ill-conditioned in fp32
        v = [smax]        # (LMS order 6 through sigmas 1.977, 1.971, 1.918 has coefficients of +-5e3: any two fp32 evaluations differ by 5e-3)
        for _ in range(m - 1):
            v.append(v[-1] / rng.uniform(1.15, 4.0))
        return v

    return torch.tensor(pick(n) + [0.0] if kind == "manual" else pick(n + 1), dtype=torch.float32)


def test_samplers_match_oracle_on_random_schedules(monkeypatch):
    _stub(monkeypatch)
    rng = random.Random(0)
    eta = lambda: rng.choice([0., 0.4, 1.])
    sn = lambda: rng.uniform(0.5, 1.2)
    cases = [("sample_euler", lambda: {}, False), ("sample_heun", lambda: {}, False), ("sample_dpmpp_2m", lambda: {}, False),
             ("sample_dpm_2", lambda: {}, False), ("sample_lms", lambda: dict(order=rng.randint(1, 6)), False),
             ("sample_euler_ancestral", lambda: dict(eta=rng.choice([0., 0.3, 1., 1.7]), s_noise=sn()), True),
             ("sample_dpm_2_ancestral", lambda: dict(eta=eta(), s_noise=sn()), True),
             ("sample_dpmpp_2s_ancestral", lambda: dict(eta=eta(), s_noise=sn()), True),
             ("sample_dpmpp_sde", lambda: dict(eta=eta(), s_noise=sn(), r=rng.choice([0.5, 0.3, 0.8])), True),
             ("sample_dpmpp_2m_sde", lambda: dict(eta=eta(), s_noise=sn(), solver_type=rng.choice(["heun", "midpoint"])), True),
             ("sample_dpmpp_3m_sde", lambda: dict(eta=eta(), s_noise=sn()), True)]
    compared, seen = 0, set()
    for trial in range(400):
        name, make_kw, noisy = rng.choice(cases)
        sig = _schedule(rng)
        x = torch.randn(2, 1, 4, 4, generator=torch.Generator().manual_seed(trial)) * float(sig[0])
        kw = make_kw()
        g = torch.Generator().manual_seed(1000 + trial)
        draws = [torch.randn(x.shape, generator=g) for _ in range(64)]

        def sampler():
            it = iter(draws)
            return lambda a, b: next(it)

        extra = lambda: dict(kw, noise_sampler=sampler()) if noisy else dict(kw)
        try:
            want = getattr(O, name)(toy2, x, sig, **extra())
        except ValueError:                                   # e.g. an LMS order the schedule is too short for: same refusal expected
            with pytest.raises(ValueError):
                getattr(S, name)(toy2, x, sig, disable=True, **extra())
            continue
        got = getattr(S, name)(toy2, x, sig, disable=True, **extra())
        scale = max(float(want.abs().max()), float(x.abs().max()), 1e-3)
        # north-star rtol on the signal's scale; nearly repeated sigmas (h -> 0 in the multistep formulas) are ill-conditioned in fp32
        assert float((got - want).abs().max()) <= 1e-3 * scale, (name, kw, sig.tolist(), float((got - want).abs().max()), scale)
        compared += 1
        seen.add(name)
    assert compared >= 380 and len(seen) == len(cases)


def test_dpm_solvers_match_oracle_on_random_arguments(monkeypatch):
    """sample_dpm_fast (every nfe 3..15, with and without ancestral noise) and sample_dpm_adaptive (orders, tolerances, PID coefficients)
    against the oracle on random sigma ranges; the adaptive runs must take the same accept / reject decisions."""
    _stub(monkeypatch)
    from k_diffusion import _native

    def dpm_error(lo, hi, prev, atol, rtol):
        delta = torch.maximum(torch.tensor(atol), torch.tensor(rtol) * torch.maximum(lo.abs(), prev.abs()))
        return float(torch.linalg.norm((lo - hi) / delta) / lo.numel() ** 0.5)

    monkeypatch.setattr(_native, "dpm_error", dpm_error)
    rng = random.Random(1)
    same_decisions = 0
    for trial in range(60):
        smin = 10 ** rng.uniform(-2.3, -0.5)
        smax = smin * 10 ** rng.uniform(1.0, 3.3)
        x = torch.randn(2, 1, 4, 4, generator=torch.Generator().manual_seed(trial)) * smax
        g = torch.Generator().manual_seed(500 + trial)
        draws = [torch.randn(x.shape, generator=g) for _ in range(400)]

        def sampler():
            it = iter(draws)
            return lambda a, b: next(it)

        eta = rng.choice([0., 0., 0.5])
        noise = lambda: dict(eta=eta, s_noise=0.9, noise_sampler=sampler()) if eta else {}
        scale = max(float(x.abs().max()), 1e-3)
        if trial % 2 == 0:
            n = rng.randint(3, 15)
            want = O.sample_dpm_fast(toy2, x, smin, smax, n, **noise())
            got = S.sample_dpm_fast(toy2, x, smin, smax, n, disable=True, **noise())
            assert float((got - want).abs().max()) <= 1e-3 * scale, ("dpm_fast", n, eta, smin, smax, float((got - want).abs().max()))
        else:
            kw = dict(order=rng.choice([2, 3]), rtol=rng.choice([0.05, 0.02]), atol=rng.choice([0.0078, 0.003]), h_init=rng.choice([0.05, 0.2]),
                      pcoeff=rng.choice([0., 0.2]), icoeff=rng.choice([1., 0.7]), dcoeff=rng.choice([0., 0.1]))
            want, winfo = O.sample_dpm_adaptive(toy2, x, smin, smax, **kw, **noise())
            got, info = S.sample_dpm_adaptive(toy2, x, smin, smax, disable=True, return_info=True, **kw, **noise())
            same_decisions += info == winfo
            if info == winfo:      # (a decision sitting exactly on the accept threshold may flip under a 1e-7 change of the error norm)
                assert float((got - want).abs().max()) <= 2e-3 * scale, ("dpm_adaptive", kw, eta, smin, smax, float((got - want).abs().max()))
    assert same_decisions >= 27          # of 30 adaptive runs


def test_schedules_and_discrete_schedule_bit_identical_on_random_arguments():
    """Host-side schedule math is the reference's own torch op sequence: bit-identical to the oracle (which the KATs pin to the reference)
    for random arguments -- the four sigma schedules, get_ancestral_step on fp32 tensors, and DiscreteSchedule's int64 / interpolated
    sigma_to_t, t_to_sigma and get_sigmas on random tables."""
    rng = random.Random(2)
    for _ in range(200):
        n = rng.randint(1, 60)
        smin = 10 ** rng.uniform(-3, 0)
        smax = smin * 10 ** rng.uniform(0.1, 4)
        rho = rng.uniform(0.5, 9)
        assert torch.equal(S.get_sigmas_karras(n, smin, smax, rho=rho), O.get_sigmas_karras(n, smin, smax, rho=rho))
        assert torch.equal(S.get_sigmas_exponential(n, smin, smax), O.get_sigmas_exponential(n, smin, smax))
        assert torch.equal(S.get_sigmas_polyexponential(n, smin, smax, rho=rho / 4), O.get_sigmas_polyexponential(n, smin, smax, rho=rho / 4))
        bd, bm, eps = rng.uniform(5, 25), rng.uniform(0.05, 0.5), 10 ** rng.uniform(-4, -2)
        assert torch.equal(S.get_sigmas_vp(n, beta_d=bd, beta_min=bm, eps_s=eps), O.get_sigmas_vp(n, beta_d=bd, beta_min=bm, eps_s=eps))
        a, b = torch.tensor(smax, dtype=torch.float32), torch.tensor(smin, dtype=torch.float32)
        for eta in (0., 0.3, 1., 2.5):
            got, want = S.get_ancestral_step(a, b, eta), O.get_ancestral_step(a, b, eta)
            assert all(float(g) == float(w) for g, w in zip(got, want)), (smax, smin, eta, got, want)
    for _ in range(40):
        m = rng.randint(2, 1000)
        lo = 10 ** rng.uniform(-3, -1)
        table = torch.exp(torch.linspace(math.log(lo), math.log(lo * 10 ** rng.uniform(1, 4)), m)
                          + 0.3 * torch.rand(m, generator=torch.Generator().manual_seed(m)).cumsum(0) / m)
        quantize = rng.random() < 0.5
        ours, ref = K.external.DiscreteSchedule(table, quantize), O.DiscreteScheduleOracle(table, quantize)
        q = torch.exp(torch.empty(64).uniform_(math.log(lo) - 1, float(table[-1].log()) + 1, generator=torch.Generator().manual_seed(m + 1)))
        for qz in (None, True, False):
            got, want = ours.sigma_to_t(q, quantize=qz), ref.sigma_to_t(q, quantize=qz)
            assert got.dtype == want.dtype and torch.equal(got, want)
        t = torch.empty(64).uniform_(0, m - 1, generator=torch.Generator().manual_seed(m + 2))
        assert torch.equal(ours.t_to_sigma(t), ref.t_to_sigma(t))
        k = rng.randint(1, 50)
        assert torch.equal(ours.get_sigmas(k), ref.get_sigmas(k)) and torch.equal(ours.get_sigmas(), ref.get_sigmas())
        assert float(ours.sigma_min) == float(table[0]) and float(ours.sigma_max) == float(table[-1])
