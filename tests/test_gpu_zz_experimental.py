"""Experimental kernels (default off).  The file name sorts after every other test module on purpose: should an experimental
kernel ever fault, the sticky CUDA error can only affect tests of this file."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]
DEV = "cuda"


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
