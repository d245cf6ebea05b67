"""Shared helpers of the checker side (tests/, __graft_entry__.smoke(), bench.py's CPU legs).  TEST INFRASTRUCTURE."""
import torch

from . import kdiff_oracle as O


def synth_sd(shapes, seed=1):
    """State dict from the key/shape/seed recipe (the one oracle/make_golden.py fed the reference)."""
    from k_diffusion import synth
    base = {}
    for k, s in shapes.items():
        if k.endswith("pos_emb.freqs"):
            base[k] = O.rope_freqs(s[1] * 8, s[0])        # freqs [nh, d_head // 8]
    return synth.synth_state_dict(shapes, seed, base)


def assert_close(a, b, rtol=1e-3, atol=1e-5, what=""):
    """north_star tolerance: rtol 1e-3 / atol 1e-5 on fp32."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    bad = err > atol + rtol * b.abs()
    if bad.any():
        i = err.argmax()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} outside rtol={rtol} atol={atol}; "
                             f"max abs err {err.max():.3e} at {int(i)} (got {a.flatten()[i]:.6e}, want {b.flatten()[i]:.6e})")
