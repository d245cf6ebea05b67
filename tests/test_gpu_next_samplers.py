"""SURVEY 8(f) row 1 on the GPU: the seven remaining fixed-schedule samplers (sample_dpm_2, sample_dpm_2_ancestral, sample_lms,
sample_dpmpp_2s_ancestral, sample_dpmpp_sde, sample_dpmpp_2m_sde, sample_dpmpp_3m_sde).

Their oracle restatements are pinned against the reference and their host plans / executor are verified on the CPU
(tests/test_oracle_golden.py, tests/test_host_logic.py).  Here the CUDA path runs (a) all 14 trajectories recorded from the
real reference (oracle/make_golden_next.py) around an opaque model, (b) every entry point on the native cfg1 model against the
CPU oracle at the north-star tolerance, (c) the CUDA-graph cache rules for noise samplers.
"""
import pytest
import torch

import k_diffusion as K
from conftest import assert_close, load_npz
from oracle import kdiff_oracle as O
from test_gpu_parity import build

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
S = K.sampling
DEV = "cuda"
toy2 = lambda x, s, **kw: x / (1 + s[:, None, None, None] ** 2) + 0.1 * torch.tanh(x)

# name in the fixture -> (entry point, kwargs, draws noise)
TRAJECTORIES = {
    "sample_dpm_2": ("sample_dpm_2", {}, False),
    "sample_lms": ("sample_lms", {}, False),
    "sample_lms_order2": ("sample_lms", dict(order=2), False),
    "sample_dpm_2_ancestral": ("sample_dpm_2_ancestral", {}, True),
    "sample_dpm_2_ancestral_eta05": ("sample_dpm_2_ancestral", dict(eta=0.5, s_noise=0.9), True),
    "sample_dpmpp_2s_ancestral": ("sample_dpmpp_2s_ancestral", {}, True),
    "sample_dpmpp_2s_ancestral_eta0": ("sample_dpmpp_2s_ancestral", dict(eta=0.), True),
    "sample_dpmpp_sde": ("sample_dpmpp_sde", {}, True),
    "sample_dpmpp_sde_r03": ("sample_dpmpp_sde", dict(eta=0.7, s_noise=0.9, r=0.3), True),
    "sample_dpmpp_2m_sde": ("sample_dpmpp_2m_sde", {}, True),
    "sample_dpmpp_2m_sde_heun": ("sample_dpmpp_2m_sde", dict(eta=0.6, solver_type="heun"), True),
    "sample_dpmpp_2m_sde_eta0": ("sample_dpmpp_2m_sde", dict(eta=0.), True),
    "sample_dpmpp_3m_sde": ("sample_dpmpp_3m_sde", {}, True),
    "sample_dpmpp_3m_sde_eta05": ("sample_dpmpp_3m_sde", dict(eta=0.5, s_noise=0.8), True),
}


@pytest.mark.parametrize("key", sorted(TRAJECTORIES))
def test_opaque_model_next_samplers_vs_reference_trajectories(key):
    z = load_npz("toy_next_samplers.npz")
    x, sig, nz = z["x"].to(DEV), z["sigmas"].to(DEV), z["noise"].to(DEV)
    fn, kw, noisy = TRAJECTORIES[key]
    if noisy:
        it = iter(nz)
        kw = dict(kw, noise_sampler=lambda a, b: next(it))
    assert_close(getattr(S, fn)(toy2, x, sig, disable=True, **kw), z[key], rtol=1e-4, atol=2e-5, what=key)


@pytest.mark.parametrize("order", [5, 6, 7, 10])
def test_sample_lms_any_order_vs_reference(order):
    """x + more than five derivative buffers: the update is a chain of lincomb launches (oracle/make_golden_lms.py recorded the reference)"""
    z = load_npz("toy_lms_high_order.npz")
    got = S.sample_lms(toy2, z["x"].to(DEV), z["sigmas"].to(DEV), disable=True, order=order)
    assert_close(got, z[f"sample_lms_order{order}"], rtol=1e-4, atol=2e-5, what=f"lms order {order}")


def test_native_model_next_samplers_vs_oracle():
    """cfg1 (MNIST transformer, class-conditional, fp32 exact path): CUDA path against the CPU oracle on the same inputs,
    rtol 1e-3 / atol 1e-5, all seven entry points."""
    cfg, sd, inner, model, z = build("cfg1_mnist")
    x, sigmas = z["x"].to(DEV), z["sigmas"].to(DEV)
    cc = z["class_cond"]
    ea = dict(class_cond=cc.to(DEV))
    oracle_model = O.make_denoiser(sd, cfg["model"])
    om = lambda xx, ss, **kw: oracle_model(xx, ss, class_cond=cc)
    noise = z["noise"]

    def pair():
        it_g, it_o = iter(noise.to(DEV)), iter(noise)
        return (lambda a, b: next(it_g)), (lambda a, b: next(it_o))

    assert_close(S.sample_dpm_2(model, x, sigmas, extra_args=ea, disable=True), O.sample_dpm_2(om, z["x"], z["sigmas"]), what="dpm_2")
    assert_close(S.sample_lms(model, x, sigmas, extra_args=ea, disable=True), O.sample_lms(om, z["x"], z["sigmas"]), what="lms")
    for name in ("sample_dpmpp_2m_sde", "sample_dpmpp_3m_sde"):
        ng, no = pair()
        got = getattr(S, name)(model, x, sigmas, extra_args=ea, disable=True, noise_sampler=ng)
        assert_close(got, getattr(O, name)(om, z["x"], z["sigmas"], noise_sampler=no), what=name)
    # two noise draws per step: the fixture records 10 tensors -> 5 Karras steps
    sig5 = K.sampling.get_sigmas_karras(5, 1e-2, 80)
    for name in ("sample_dpm_2_ancestral", "sample_dpmpp_2s_ancestral", "sample_dpmpp_sde"):
        ng, no = pair()
        got = getattr(S, name)(model, x, sig5.to(DEV), extra_args=ea, disable=True, noise_sampler=ng)
        assert_close(got, getattr(O, name)(om, z["x"], sig5, noise_sampler=no), what=name)


def test_brownian_default_graph_equals_eager(monkeypatch):
    """Unconditional native model + an explicit Brownian-tree noise sampler: the captured graph must replay the eager launches."""
    cfg, sd, inner, model, z = build("sw64")
    x, sigmas = z["x"].to(DEV), z["sigmas"].to(DEV)
    ns = S.BrownianTreeNoiseSampler(x, float(sigmas[sigmas > 0].min()), float(sigmas.max()), seed=list(range(x.shape[0])))
    S.clear_graph_cache()
    g1 = S.sample_dpmpp_2m_sde(model, x, sigmas, disable=True, noise_sampler=ns)
    monkeypatch.setenv("KDB200_CUDA_GRAPH", "0")
    e1 = S.sample_dpmpp_2m_sde(model, x, sigmas, disable=True, noise_sampler=ns)
    assert torch.isfinite(g1).all() and torch.equal(g1, e1)


def test_graph_cache_is_keyed_on_noise_content_not_object_identity(monkeypatch):
    """a graph must never replay a freed sampler's seeds.  Back-to-back default samplers (fresh random seed,
    freed on return, CPython reuses the address) must differ; two samplers with the same explicit seeds must agree; a new seed
    list must give the eager result for THAT list."""
    cfg, sd, inner, model, z = build("sw64")
    x, sigmas = z["x"].to(DEV), z["sigmas"].to(DEV)
    B = x.shape[0]
    lo, hi = float(sigmas[sigmas > 0].min()), float(sigmas.max())
    S.clear_graph_cache()
    torch.manual_seed(0)
    a = S.sample_dpmpp_2m_sde(model, x, sigmas, disable=True)
    b = S.sample_dpmpp_2m_sde(model, x, sigmas, disable=True)
    assert not torch.equal(a, b), "two default (random-seed) Brownian samplers replayed the same noise"
    mk = lambda seeds: S.BrownianTreeNoiseSampler(x, lo, hi, seed=seeds)
    s1 = S.sample_euler_ancestral(model, x, sigmas, disable=True, noise_sampler=mk(list(range(10, 10 + B))))
    s2 = S.sample_euler_ancestral(model, x, sigmas, disable=True, noise_sampler=mk(list(range(10, 10 + B))))
    s3 = S.sample_euler_ancestral(model, x, sigmas, disable=True, noise_sampler=mk(list(range(50, 50 + B))))
    assert torch.equal(s1, s2) and not torch.equal(s1, s3)
    monkeypatch.setenv("KDB200_CUDA_GRAPH", "0")
    e3 = S.sample_euler_ancestral(model, x, sigmas, disable=True, noise_sampler=mk(list(range(50, 50 + B))))
    assert torch.equal(s3, e3)


def test_graph_cache_is_keyed_on_sigma_data():
    """sigma_data is baked into the captured patch-in / patch-out kernels."""
    cfg, sd, inner, model, z = build("sw64")
    x, sigmas = z["x"].to(DEV), z["sigmas"].to(DEV)
    S.clear_graph_cache()
    a = S.sample_heun(model, x, sigmas, disable=True)
    other = K.Denoiser(inner, sigma_data=float(model.sigma_data) * 2)
    b = S.sample_heun(other, x, sigmas, disable=True)
    assert not torch.equal(a, b)
    want = O.sample_heun(O.make_denoiser(sd, dict(cfg["model"], sigma_data=float(model.sigma_data) * 2)), z["x"], z["sigmas"])
    assert_close(b, want, what="heun with the second Denoiser's sigma_data")


def test_class_conditional_graph_equals_eager(monkeypatch):
    """Per-sample conditioning (class labels) is captured too: the graph reads static copies of the extra_args tensors that are
    refreshed before every replay.  cfg1, fp32 exact path: graph == eager bit for bit, and new labels give new results."""
    cfg, sd, inner, model, z = build("cfg1_mnist")
    x, sigmas = z["x"].to(DEV), z["sigmas"].to(DEV)
    cc = z["class_cond"].to(DEV)
    S.clear_graph_cache()
    g1 = S.sample_heun(model, x, sigmas, extra_args=dict(class_cond=cc), disable=True)
    assert_close(g1, z["heun"], what="heun (class-conditional graph)")
    cc2 = torch.tensor([5, 5, 0, 9], device=DEV)
    g2 = S.sample_heun(model, x, sigmas, extra_args=dict(class_cond=cc2), disable=True)          # replay with refreshed labels
    monkeypatch.setenv("KDB200_CUDA_GRAPH", "0")
    e1 = S.sample_heun(model, x, sigmas, extra_args=dict(class_cond=cc), disable=True)
    e2 = S.sample_heun(model, x, sigmas, extra_args=dict(class_cond=cc2), disable=True)
    assert torch.equal(g1, e1) and torch.equal(g2, e2) and not torch.equal(g1, g2)
    with pytest.raises(IndexError, match="class_cond"):
        S.sample_heun(model, x, sigmas, extra_args=dict(class_cond=torch.tensor([0, 1, 2, 11], device=DEV)), disable=True)


# ------------------------------------------------------------------------------------------
# classifier-free guidance wrapper (reference train.py:333-344)
# ------------------------------------------------------------------------------------------

def test_cfg_wrapper_vs_oracle(monkeypatch):
    """cfg1 (MNIST transformer, 10 classes + unconditional token, fp32 exact path): make_cfg_model_fn on the native model through
    sample_dpmpp_2m_sde(eta=0, solver_type='heun') -- the sampler / settings of the reference's demo() -- against the oracle's
    restatement of the closure, rtol 1e-3 / atol 1e-5; graph == eager; the wrapper called directly; cfg_scale == 1 returns the model."""
    cfg, sd, inner, model, z = build("cfg1_mnist")
    x, sigmas = z["x"].to(DEV), z["sigmas"].to(DEV)
    cc = torch.tensor([0, 3, 7, 9])
    oracle_model = O.make_denoiser(sd, cfg["model"])
    o_fn = O.make_cfg_model_fn(oracle_model, 3.0, 10)
    want = O.sample_dpmpp_2m_sde(o_fn, z["x"], z["sigmas"], noise_sampler=None, extra_args=dict(class_cond=cc), eta=0.0, solver_type="heun")
    fn = S.make_cfg_model_fn(model, 3.0, 10)
    assert S.make_cfg_model_fn(model, 1.0, 10) is model
    S.clear_graph_cache()
    got = S.sample_dpmpp_2m_sde(fn, x, sigmas, extra_args=dict(class_cond=cc.to(DEV)), eta=0.0, solver_type="heun", disable=True)
    assert_close(got, want, what="cfg dpmpp_2m_sde (graph)")
    ref = load_npz("cfg1_cfg.npz")            # recorded from the reference's own closure (oracle/make_golden_cfg.py), same inputs
    assert torch.equal(ref["class_cond"], cc)
    assert_close(got, ref["dpmpp_2m_sde_heun_eta0"], what="cfg dpmpp_2m_sde vs the reference closure")
    assert_close(fn(x, ref["sigma"].to(DEV), class_cond=cc.to(DEV)), ref["model_fn"], what="cfg model_fn vs the reference closure")
    # direct call == oracle closure on one evaluation
    sig = torch.tensor([0.3, 1.0, 5.0, 40.0])
    assert_close(fn(x, sig.to(DEV), class_cond=cc.to(DEV)), o_fn(z["x"], sig, class_cond=cc), what="cfg model_fn")
    # new labels through the cached graph, against eager launches
    cc2 = torch.tensor([1, 1, 2, 8], device=DEV)
    g2 = S.sample_dpmpp_2m_sde(fn, x, sigmas, extra_args=dict(class_cond=cc2), eta=0.0, solver_type="heun", disable=True)
    monkeypatch.setenv("KDB200_CUDA_GRAPH", "0")
    e1 = S.sample_dpmpp_2m_sde(fn, x, sigmas, extra_args=dict(class_cond=cc.to(DEV)), eta=0.0, solver_type="heun", disable=True)
    e2 = S.sample_dpmpp_2m_sde(fn, x, sigmas, extra_args=dict(class_cond=cc2), eta=0.0, solver_type="heun", disable=True)
    assert torch.equal(got, e1) and torch.equal(g2, e2) and not torch.equal(got, g2)
    # opaque models work through the same wrapper
    toy = lambda xx, ss, class_cond: xx / (1 + ss[:, None, None, None] ** 2) * (1 + 0.1 * class_cond[:, None, None, None].float())
    tfn = S.make_cfg_model_fn(toy, 2.0, 10)
    xin, uncond = x, torch.full_like(cc2, 10)
    ref = toy(xin, sig.to(DEV), uncond) + (toy(xin, sig.to(DEV), cc2) - toy(xin, sig.to(DEV), uncond)) * 2.0
    assert_close(tfn(xin, sig.to(DEV), class_cond=cc2), ref, rtol=1e-5, atol=1e-6, what="cfg around an opaque model")


# ------------------------------------------------------------------------------------------------------------------------------
# DPM-Solver fast / adaptive (reference sampling.py:303-516; goldens from oracle/make_golden_dpm.py)
# ------------------------------------------------------------------------------------------------------------------------------
DPM_ADAPTIVE = {"dpm_adaptive_o3": dict(), "dpm_adaptive_o2": dict(order=2), "dpm_adaptive_o3_tight": dict(rtol=0.01, atol=0.002, h_init=0.1),
                "dpm_adaptive_o3_pid": dict(pcoeff=0.2, icoeff=0.7, dcoeff=0.1, accept_safety=0.9), "dpm_adaptive_o3_eta05": dict(eta=0.5, s_noise=0.9)}


def _scale_close(got, want, what, rel=1e-3):
    """North-star rtol against the signal's scale: the sigma-80 toy problem amplifies a 1e-7 coefficient change to ~2e-4 (see the
    comment in tests/test_host_logic.py::test_dpm_solver_plans_and_entry_points_with_stubbed_kernels)."""
    err = float((got.detach().float().cpu() - want).abs().max())
    assert err <= rel * float(want.abs().max()), f"{what}: max abs err {err:.3e} vs scale {float(want.abs().max()):.3e}"


@pytest.mark.parametrize("n", [4, 5, 6, 9, 10])
def test_dpm_fast_opaque_model_vs_reference(n):
    z = load_npz("toy_dpm_solvers.npz")
    _scale_close(S.sample_dpm_fast(toy2, z["x"].to(DEV), 1e-2, 80., n, disable=True), z[f"dpm_fast_n{n}"], f"dpm_fast n={n}", rel=2e-4)


def test_dpm_fast_stochastic_opaque_model_vs_reference():
    z = load_npz("toy_dpm_solvers.npz")
    x, nz = z["x"].to(DEV), z["noise"].to(DEV)
    it = iter(nz)
    got = S.sample_dpm_fast(toy2, x, 1e-2, 80., 7, disable=True, eta=0.5, s_noise=0.9, noise_sampler=lambda a, b: next(it))
    _scale_close(got, z["dpm_fast_n7_eta05"], "dpm_fast n=7 eta=0.5", rel=2e-4)
    it = iter(nz)
    _scale_close(S.sample_dpm_fast(toy2, x, 1e-2, 80., 6, disable=True, eta=1.0, noise_sampler=lambda a, b: next(it)), z["dpm_fast_n6_eta1"], "dpm_fast n=6 eta=1")


@pytest.mark.parametrize("name", sorted(DPM_ADAPTIVE))
def test_dpm_adaptive_opaque_model_vs_reference(name):
    """Same accept / reject sequence as the reference (the error norm is kdb_solver_dpm_error) and the same samples."""
    z = load_npz("toy_dpm_solvers.npz")
    kw = dict(DPM_ADAPTIVE[name])
    if kw.get("eta"):
        it = iter(z["noise"].to(DEV))
        kw["noise_sampler"] = lambda a, b: next(it)
    got, info = S.sample_dpm_adaptive(toy2, z["x"].to(DEV), 1e-2, 80., disable=True, return_info=True, **kw)
    assert [info[k] for k in ("steps", "nfe", "n_accept", "n_reject")] == [int(v) for v in z[name + "_info"]], (name, info)
    _scale_close(got, z[name], name)


def test_dpm_error_kernel_matches_torch():
    from k_diffusion import _native
    g = torch.Generator(device=DEV).manual_seed(5)
    lo = torch.randn(4, 3, 64, 64, device=DEV, generator=g) * 3
    hi = lo + torch.randn(4, 3, 64, 64, device=DEV, generator=g) * 0.05
    prev = torch.randn(4, 3, 64, 64, device=DEV, generator=g) * 3
    delta = torch.maximum(torch.tensor(0.0078, device=DEV), 0.05 * torch.maximum(lo.abs(), prev.abs()))
    want = float(torch.linalg.norm(((lo - hi) / delta).double()) / lo.numel() ** 0.5)
    got = _native.dpm_error(lo, hi, prev, 0.0078, 0.05)
    assert abs(got - want) <= 1e-5 * want
    assert _native.dpm_error(lo, hi, prev, 0.0078, 0.05) == got          # deterministic reduction


def test_dpm_solvers_native_model_vs_oracle():
    """cfg1 (MNIST transformer, class-conditional, fp32 exact path): the native engine behind sample_dpm_fast (graph-captured op plan)
    and sample_dpm_adaptive (eager host loop) against the CPU oracle on the same inputs."""
    cfg, sd, inner, model, z = build("cfg1_mnist")
    x = z["x"].to(DEV)
    cc = z["class_cond"]
    ea = dict(class_cond=cc.to(DEV))
    oracle_model = O.make_denoiser(sd, cfg["model"])
    om = lambda xx, ss, **kw: oracle_model(xx, ss, class_cond=cc)
    _scale_close(S.sample_dpm_fast(model, x, 1e-2, 80., 7, extra_args=ea, disable=True), O.sample_dpm_fast(om, z["x"], 1e-2, 80., 7), "native dpm_fast")
    got, info = S.sample_dpm_adaptive(model, x, 1e-2, 80., extra_args=ea, disable=True, return_info=True)
    want, winfo = O.sample_dpm_adaptive(om, z["x"], 1e-2, 80.)
    assert info == winfo, (info, winfo)
    _scale_close(got, want, "native dpm_adaptive", rel=2e-3)
