"""Pin oracle/kdiff_oracle.py against fixtures produced by the real reference (oracle/make_golden.py).

CPU only.  Bit-exact where the reference arithmetic is reproduced op for op (schedules, integer
schedule indices, position grids, RoPE frequencies, masks); rtol 1e-3 / atol 1e-5 for model outputs
and sampler trajectories (the oracle uses explicit softmax instead of SDPA's fused CPU kernel).
"""
import numpy as np
import torch

from conftest import assert_close, bits, load_fixture, load_npz, synth_sd, unhex
from oracle import kdiff_oracle as O


def test_schedules_bit_exact(kat):
    fns = dict(karras=O.get_sigmas_karras, exponential=O.get_sigmas_exponential,
               polyexponential=O.get_sigmas_polyexponential, vp=O.get_sigmas_vp)
    assert len(kat["schedules"]) >= 10
    for e in kat["schedules"]:
        assert bits(fns[e["fn"]](*e["args"])) == e["hex"], e


def test_survey_karras_hex():
    # SURVEY.md section 8c known-answer vector, recorded independently of make_golden.py
    want = "429ffffe 42320d70 41bbcc01 4139b71d 40a9ba6e 400c86a7 3f4cdeb4 3e7bde3d 3d739676 3c23d706 00000000".split()
    assert bits(O.get_sigmas_karras(10, 1e-2, 80)) == want


def test_ancestral_step(kat):
    for e in kat["ancestral"]:
        d, u = O.get_ancestral_step(unhex([e["sigma_from"]])[0], unhex([e["sigma_to"]])[0], eta=e["eta"])
        assert float(d) == e["down"] and float(u) == e["up"], e


def test_toy_sampler_kats(kat):
    sig = O.get_sigmas_karras(5, 1e-2, 80)
    toy = lambda x, s, **kw: 0.5 * x
    x = torch.ones(2, 1, 4, 4)
    for name in ("sample_euler", "sample_heun", "sample_dpmpp_2m"):
        got = float(getattr(O, name)(toy, x, sig).flatten()[0])
        assert got == kat["toy"][name], name
    # SURVEY.md section 8c values
    assert abs(kat["toy"]["sample_heun"] - 0.02896302566) < 1e-9
    assert abs(kat["toy"]["sample_dpmpp_2m"] + 0.03126364946) < 1e-9


def test_nonlinear_toy_samplers_bit_exact():
    z = load_npz("toy_samplers.npz")
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    for name in ("sample_euler", "sample_heun", "sample_dpmpp_2m"):
        assert torch.equal(getattr(O, name)(toy2, z["x"], z["sigmas"]), z[name]), name
    it = iter(z["noise"])
    assert torch.equal(O.sample_euler_ancestral(toy2, z["x"], z["sigmas"], noise_sampler=lambda a, b: next(it)), z["sample_euler_ancestral"])
    it = iter(z["noise"])
    got = O.sample_euler_ancestral(toy2, z["x"], z["sigmas"], eta=0.5, s_noise=0.9, noise_sampler=lambda a, b: next(it))
    assert torch.equal(got, z["sample_euler_ancestral_eta05"])


def test_next_row_samplers_against_reference():
    """SURVEY 8(f) row 1 (not yet on the CUDA path): oracle restatements pinned against reference trajectories."""
    z = load_npz("toy_next_samplers.npz")
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    x, sig = z["x"], z["sigmas"]

    def ns():
        it = iter(z["noise"])
        return lambda a, b: next(it)

    cases = {
        "sample_dpm_2": lambda: O.sample_dpm_2(toy2, x, sig),
        "sample_dpm_2_ancestral": lambda: O.sample_dpm_2_ancestral(toy2, x, sig, noise_sampler=ns()),
        "sample_dpm_2_ancestral_eta05": lambda: O.sample_dpm_2_ancestral(toy2, x, sig, eta=0.5, s_noise=0.9, noise_sampler=ns()),
        "sample_dpmpp_2s_ancestral": lambda: O.sample_dpmpp_2s_ancestral(toy2, x, sig, noise_sampler=ns()),
        "sample_dpmpp_2s_ancestral_eta0": lambda: O.sample_dpmpp_2s_ancestral(toy2, x, sig, eta=0.0, noise_sampler=ns()),
        "sample_dpmpp_sde": lambda: O.sample_dpmpp_sde(toy2, x, sig, ns()),
        "sample_dpmpp_sde_r03": lambda: O.sample_dpmpp_sde(toy2, x, sig, ns(), eta=0.7, s_noise=0.9, r=0.3),
        "sample_dpmpp_2m_sde": lambda: O.sample_dpmpp_2m_sde(toy2, x, sig, ns()),
        "sample_dpmpp_2m_sde_heun": lambda: O.sample_dpmpp_2m_sde(toy2, x, sig, ns(), eta=0.6, solver_type="heun"),
        "sample_dpmpp_2m_sde_eta0": lambda: O.sample_dpmpp_2m_sde(toy2, x, sig, ns(), eta=0.0),
        "sample_dpmpp_3m_sde": lambda: O.sample_dpmpp_3m_sde(toy2, x, sig, ns()),
        "sample_dpmpp_3m_sde_eta05": lambda: O.sample_dpmpp_3m_sde(toy2, x, sig, ns(), eta=0.5, s_noise=0.8),
    }
    for name, run in cases.items():       # same torch op sequence as the reference on the same CPU -> identical bits
        assert torch.equal(run(), z[name]), name
    # linear multistep: the reference integrates the Lagrange basis with scipy quad (epsrel 1e-4), the oracle exactly
    assert_close(O.sample_lms(toy2, x, sig), z["sample_lms"], rtol=1e-5, atol=1e-5, what="sample_lms")
    assert_close(O.sample_lms(toy2, x, sig, order=2), z["sample_lms_order2"], rtol=1e-5, atol=1e-5, what="sample_lms order 2")
    import pytest
    with pytest.raises(ValueError):
        O.lms_coefficient(3, sig.numpy(), 1, 0)
    with pytest.raises(ValueError):
        O.sample_dpmpp_2m_sde(toy2, x, sig, ns(), solver_type="euler")


def _discrete():
    betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2
    ac = torch.cumprod(1 - betas, 0)
    return O.DiscreteScheduleOracle(((1 - ac) / ac) ** 0.5, True)


def test_discrete_schedule(kat):
    d, ds = kat["discrete"], _discrete()
    assert bits(ds.get_sigmas(10)) == d["get_sigmas_10"]
    assert bits(ds.get_sigmas(37)) == d["get_sigmas_37"]
    full = ds.get_sigmas()
    assert len(full) == d["get_sigmas_all_len"] == 1001
    assert bits(full[:5]) == d["get_sigmas_all_head"] and bits(full[-5:]) == d["get_sigmas_all_tail"]
    t = ds.sigma_to_t(ds.get_sigmas(10)[:-1])
    assert t.dtype == torch.int64 and t.tolist() == d["roundtrip_t"] == [999, 888, 777, 666, 555, 444, 333, 222, 111, 0]
    q = unhex(d["query"])
    assert ds.sigma_to_t(q).tolist() == d["t_quant"]
    assert bits(ds.sigma_to_t(q, quantize=False)) == d["t_interp"]
    assert bits(ds.t_to_sigma(torch.tensor([0.0, 0.5, 17.25, 998.9, 999.0]))) == d["t_to_sigma"]


def test_scalings(kat):
    s = kat["scalings"]
    cs, co, ci = O.karras_scalings(torch.tensor(s["sigma"]), s["sigma_data"])
    assert bits(cs) == s["c_skip"] and bits(co) == s["c_out"] and bits(ci) == s["c_in"]


def test_layer_kats(kat):
    L = kat["layers"]
    assert bits(O.rope_freqs(64, 2)) == L["rope_freqs_32_2"]
    assert bits(O.rope_freqs(64, 8)) == L["rope_freqs_32_8"]
    assert bits(O.make_axial_pos(4, 4)) == L["axial_pos_4_4"]
    assert bits(O.make_axial_pos(7, 7)) == L["axial_pos_7_7"]
    assert bits(O.make_axial_pos(4, 8)) == L["axial_pos_4_8"]
    assert bits(O.downscale_pos(O.make_axial_pos(8, 8))) == L["downscale_pos_8_8"]
    m = O.shifted_window_allow(2, 3, 4, 2)
    assert "".join("1" if b else "0" for b in m.flatten().tolist()) == L["sw_mask_2_3_4_4_2"]
    a = torch.arange(16.0).view(1, 4, 4, 1)
    assert O.token_merge(a, torch.eye(4), 2, 2).flatten().tolist() == L["token_merge_4x4"]


def test_cfg1_mnist_forward_and_samplers():
    cfg, shapes, z = load_fixture("cfg1_mnist")
    sd, mcfg = synth_sd(shapes), cfg["model"]
    x, cc = z["x"], z["class_cond"]
    assert_close(O.model_forward(sd, mcfg, x * 0.01, z["sigma"], class_cond=cc), z["inner"], what="inner")
    assert_close(O.model_forward(sd, mcfg, x * 0.01, z["sigma"], aug_cond=z["aug_cond"], class_cond=cc), z["inner_aug"], what="inner_aug")
    model = O.make_denoiser(sd, mcfg)
    assert_close(model(x, z["sigma"], class_cond=cc), z["denoised"], what="denoised")
    ea = dict(class_cond=cc)
    assert_close(O.sample_heun(model, x, z["sigmas"], ea), z["heun"], what="heun")
    assert_close(O.sample_dpmpp_2m(model, x, z["sigmas"], ea), z["dpmpp_2m"], what="dpmpp_2m")
    assert_close(O.sample_euler(model, x, z["sigmas"], ea), z["euler"], what="euler")
    it = iter(z["noise"])
    assert_close(O.sample_euler_ancestral(model, x, z["sigmas"], ea, noise_sampler=lambda a, b: next(it)), z["euler_ancestral"], what="euler_a")
    assert z["inner"].abs().max() > 1e-3      # not the vacuous all-zero model


def test_sw64_forward_and_samplers():
    cfg, shapes, z = load_fixture("sw64")
    sd, mcfg = synth_sd(shapes), cfg["model"]
    assert_close(O.model_forward(sd, mcfg, z["x"] * 0.01, z["sigma"]), z["inner"], what="inner")
    model = O.make_denoiser(sd, mcfg)
    assert_close(model(z["x"], z["sigma"]), z["denoised"], what="denoised")
    assert_close(O.sample_heun(model, z["x"], z["sigmas"]), z["heun"], what="heun")
    assert_close(O.sample_dpmpp_2m(model, z["x"], z["sigmas"]), z["dpmpp_2m"], what="dpmpp_2m")


def test_cfg2_sw256_forward():
    cfg, shapes, z = load_fixture("cfg2_sw256")
    sd, mcfg = synth_sd(shapes), cfg["model"]
    g = torch.Generator().manual_seed(int(z["seed"]))
    x = torch.randn(1, 3, 256, 256, generator=g) * 160
    o = O.model_forward(sd, mcfg, x * 0.01, z["sigma"])
    assert_close(o[..., ::4, ::4], z["inner_sub"], what="inner_sub")
    assert abs(o.double().mean().item() - float(z["inner_mean"])) < 1e-6
    assert abs(o.double().pow(2).mean().item() / float(z["inner_sqmean"]) - 1) < 1e-4
    assert_close(O.make_denoiser(sd, mcfg)(x, z["sigma"])[..., ::4, ::4], z["denoised_sub"], what="denoised_sub")


def test_macs_match_reference_counter():
    import json
    from conftest import GOLDEN
    macs = json.loads((GOLDEN / "macs.json").read_text())
    cfg1 = json.loads((GOLDEN / "cfg1_mnist_shapes.json").read_text())["config"]["model"]
    cfg2 = json.loads((GOLDEN / "cfg2_sw256_shapes.json").read_text())["config"]["model"]
    assert O.model_macs(cfg1) == macs["cfg1_mnist"] == 346566656
    assert O.model_macs(cfg2) == macs["cfg2_sw256"] == 11730419712


def test_neighborhood_definition_properties():
    """natten is absent (parity unpinned): check the restated definition's invariants instead."""
    allow = O.neighborhood_allow(9, 12, 7)
    assert allow.shape == (108, 108) and bool((allow.sum(-1) == 49).all())
    a = allow.view(9, 12, 9, 12)
    assert bool(a[4, 6, 1:8, 3:10].all()) and not bool(a[4, 6, 0].any())              # interior: centred
    assert bool(a[0, 0, 0:7, 0:7].all()) and int(a[0, 0].sum()) == 49                 # corner: clamped inward
    assert bool(a[8, 11, 2:9, 5:12].all())
    # kernel covering the whole grid == global attention
    g = torch.Generator().manual_seed(0)
    q, k, v = (torch.randn(2, 7, 7, 2, 8, generator=g) for _ in range(3))
    assert_close(O.neighborhood_attention(q, k, v, 7), O.global_attention(q, k, v), what="na==global")


def test_cfg_wrapper_against_reference_closure():
    """O.make_cfg_model_fn against outputs of the reference's own closure (train.py:333-344, executed from the
    reference source by oracle/make_golden_cfg.py) on the cfg1 model: one call and the demo() sampler recipe."""
    cfg, shapes, z = load_fixture("cfg1_mnist")
    c = load_npz("cfg1_cfg.npz")
    model = O.make_denoiser(synth_sd(shapes, 1), cfg["model"])
    fn = O.make_cfg_model_fn(model, float(c["cfg_scale"]), int(c["num_classes"]))
    assert O.make_cfg_model_fn(model, 1.0, 10) is model
    assert_close(fn(z["x"], c["sigma"], class_cond=c["class_cond"]), c["model_fn"], what="cfg model_fn")
    got = O.sample_dpmpp_2m_sde(fn, z["x"], z["sigmas"], noise_sampler=None, extra_args=dict(class_cond=c["class_cond"]), eta=0.0, solver_type="heun")
    assert_close(got, c["dpmpp_2m_sde_heun_eta0"], what="cfg dpmpp_2m_sde")


DPM_ADAPTIVE_CASES = {"dpm_adaptive_o3": dict(), "dpm_adaptive_o2": dict(order=2), "dpm_adaptive_o3_tight": dict(rtol=0.01, atol=0.002, h_init=0.1),
                      "dpm_adaptive_o3_pid": dict(pcoeff=0.2, icoeff=0.7, dcoeff=0.1, accept_safety=0.9), "dpm_adaptive_o3_eta05": dict(eta=0.5, s_noise=0.9)}


def test_dpm_solver_fast_and_adaptive_against_reference():
    """DPM-Solver fast (every order pattern) and adaptive 12 / 23 with the PID controller, bit-identical to
    the outputs oracle/make_golden_dpm.py recorded from the reference (sampling.py:303-516), step / rejection counts included."""
    z = load_npz("toy_dpm_solvers.npz")
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    x = z["x"]

    def ns():
        it = iter(z["noise"])
        return lambda a, b: next(it)

    for n in (4, 5, 6, 9, 10):
        assert torch.equal(O.sample_dpm_fast(toy2, x, 1e-2, 80., n), z[f"dpm_fast_n{n}"]), n
    assert torch.equal(O.sample_dpm_fast(toy2, x, 1e-2, 80., 7, eta=0.5, s_noise=0.9, noise_sampler=ns()), z["dpm_fast_n7_eta05"])
    assert torch.equal(O.sample_dpm_fast(toy2, x, 1e-2, 80., 6, eta=1.0, noise_sampler=ns()), z["dpm_fast_n6_eta1"])
    for name, kw in DPM_ADAPTIVE_CASES.items():
        got, info = O.sample_dpm_adaptive(toy2, x, 1e-2, 80., noise_sampler=ns() if kw.get("eta") else None, **kw)
        assert torch.equal(got, z[name]), name
        assert [info[k] for k in ("steps", "nfe", "n_accept", "n_reject")] == [int(v) for v in z[name + "_info"]], name


def test_log_likelihood_against_reference_closed_form_and_scipy():
    """SURVEY 8(f) row 4, log_likelihood (sampling.py:280-301).  torchdiffeq (the reference's integrator) is absent, so:
    (1) the oracle equals the REFERENCE function run with the oracle's dopri5 installed as `sampling.odeint`
        (oracle/make_golden_ll.py) -- pins the ODE right-hand side, the prior term and the assembly;
    (2) for Gaussian data N(0, s^2 I) everything is analytic: log p_sigma_min(x) = sum log N(x_i; 0, s^2 + sigma_min^2)
        -- pins the integrated value (the residual at tight tolerance is the prior mismatch s^2 / sigma_max^2);
    (3) the integrator alone against scipy's independent Dormand-Prince (RK45) on a nonlinear system;
    (4) the tableau satisfies the order conditions (5th-order solution, 4th-order embedded and mid-point weights)."""
    import math
    z = load_npz("toy_log_likelihood.npz")
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    gauss = lambda x, s, **kw: x * (0.49 / (0.49 + s[:, None, None, None] ** 2))
    x = z["x"]
    for name, model, kw in (("toy", toy2, {}), ("gauss", gauss, {}), ("toy_tight", toy2, dict(atol=1e-6, rtol=1e-6))):
        ll, info = O.log_likelihood(model, x, 1e-2, 80., v=z[name + "_v"], **kw)
        assert info["fevals"] == int(z[name + "_fevals"]), name
        assert torch.equal(ll, z[name + "_ll"]), name
    # (2) closed form, float64 so that only the method's error is left
    xd = x.double()
    want = torch.distributions.Normal(0, math.sqrt(0.49 + 1e-4)).log_prob(xd).flatten(1).sum(1)
    ll, _ = O.log_likelihood(gauss, xd, 1e-2, 80., v=z["gauss_v"].double())
    assert float((ll - want).abs().max()) < 5e-4 * float(want.abs().max())            # default tolerances: a few rtol |ll|
    ll, _ = O.log_likelihood(gauss, xd, 1e-2, 800., atol=1e-9, rtol=1e-9, v=z["gauss_v"].double())
    assert float((ll - want).abs().max()) < 1e-4
    # (3) integrator vs scipy RK45 on a stiff-ish nonlinear pair
    from scipy.integrate import solve_ivp
    f_np = lambda t, y: [-y[0] * y[1] + math.sin(t), y[0] ** 2 - 0.5 * y[1]]
    f_t = lambda t, y: (-y[0] * y[1] + math.sin(t), y[0] ** 2 - 0.5 * y[1])
    y0 = (torch.tensor([1.0], dtype=torch.float64), torch.tensor([0.5], dtype=torch.float64))
    (a, b), st = O.odeint_dopri5(f_t, y0, 0.0, 7.0, 1e-10, 1e-10)
    ref = solve_ivp(f_np, (0.0, 7.0), [1.0, 0.5], method="RK45", atol=1e-12, rtol=1e-12).y[:, -1]
    assert abs(float(a) - ref[0]) < 1e-8 and abs(float(b) - ref[1]) < 1e-8 and st["n_accept"] > 20
    (a4, b4), st4 = O.odeint_dopri5(f_t, y0, 0.0, 7.0, 1e-4, 1e-4)
    assert abs(float(a4) - ref[0]) < 2e-3 and st4["n_accept"] < st["n_accept"] / 4       # the tolerance steers the step count
    # (4) order conditions
    c = np.array([0.0, *O.DOPRI5_ALPHA])
    A = np.zeros((7, 7))
    for i, row in enumerate(O.DOPRI5_BETA):
        A[i + 1, :len(row)] = row
    b, e, m = (np.array(t) for t in (O.DOPRI5_C_SOL, O.DOPRI5_C_ERR, O.DOPRI5_C_MID))
    assert np.abs(A.sum(1) - c).max() < 1e-15 and np.abs(A[6] - b).max() == 0           # first-same-as-last
    for q in range(5):
        assert abs(b @ c ** q - 1 / (q + 1)) < 1e-15
    for q in range(4):
        assert abs(e @ c ** q) < 1e-15 and abs(m @ c ** q - 0.5 ** (q + 1) / (q + 1)) < 1e-15
    for got, want_ in ((b @ A @ c, 1 / 6), (b @ A @ c ** 2, 1 / 12), (b @ (c * (A @ c)), 1 / 8), (b @ A @ A @ c, 1 / 24),
                       (b @ A @ c ** 3, 1 / 20), (b @ A @ A @ A @ c, 1 / 120), (e @ A @ c, 0), (e @ A @ A @ c, 0),
                       (m @ A @ c, 0.5 ** 3 / 6), (m @ A @ A @ c, 0.5 ** 4 / 24)):
        assert abs(got - want_) < 1e-15


def test_sample_lms_any_order_against_reference():
    """sample_lms accepts any order (sampling.py:260-277); orders 5, 6, 7 and 10 recorded from the reference (oracle/make_golden_lms.py).
    The oracle's Gauss-Legendre coefficients replace scipy quad (epsrel 1e-4) and land within 1e-6 of its outputs."""
    z = load_npz("toy_lms_high_order.npz")
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    for o in (5, 6, 7, 10):
        assert_close(O.sample_lms(toy2, z["x"], z["sigmas"], order=o), z[f"sample_lms_order{o}"], rtol=1e-5, atol=2e-6, what=f"lms order {o}")


CHURN_CASES = {"euler_churn20": ("sample_euler", dict(s_churn=20.)),
               "heun_churn3_window": ("sample_heun", dict(s_churn=3., s_tmin=0.1, s_tmax=30., s_noise=1.1)),
               "dpm_2_churn2": ("sample_dpm_2", dict(s_churn=2.))}


def test_karras_churn_against_reference(monkeypatch):
    """s_churn > 0 (sampling.py:121-127, :162-169, :192-198): gamma, sigma_hat and the injected noise, with the reference's own
    per-step randn_like draws replayed (oracle/make_golden_churn.py).  The oracle draws on every step like the reference."""
    z = load_npz("toy_churn.npz")
    toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)
    for name, (fn, kw) in CHURN_CASES.items():
        it = iter(z[name + "_eps"])
        monkeypatch.setattr(torch, "randn_like", lambda t, *a, **k: next(it))
        got = getattr(O, fn)(toy2, z["x"], z["sigmas"], **kw)
        assert torch.equal(got, z[name]), (name, float((got - z[name]).abs().max()))
