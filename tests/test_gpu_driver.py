"""checkpoint I/O + data-parallel sampling driver (reference sample.py:37-66, evaluation.py:80-90,
convert_for_inference.py:39-45, config.py:113-115) on the GPU: a synthetic inference checkpoint (safetensors with the config in its
metadata, fp16 weights like the reference's converter writes) goes through `k-diffusion_b200/sample.py` and comes out as PNGs."""
import json
import runpy
import sys

import pytest
import torch

import k_diffusion as K
from conftest import ROOT, assert_close, load_fixture, synth_sd

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
S = K.sampling
DEV = "cuda"


def test_sample_script_roundtrip(tmp_path, monkeypatch):
    import safetensors.torch as safetorch
    from PIL import Image
    cfg, shapes, _ = load_fixture("sw64")
    sd = {k: v.half() if v.dtype == torch.float32 and not k.endswith("freqs") else v for k, v in synth_sd(shapes, 1).items()}
    ckpt = tmp_path / "model.safetensors"
    safetorch.save_file(sd, str(ckpt), metadata={"config": json.dumps(cfg, indent=4)})          # convert_for_inference.py:42-45
    # config comes out of the checkpoint metadata (config.py:113-115)
    assert K.config.load_config(ckpt)["model"]["widths"] == cfg["model"]["widths"]
    prefix = tmp_path / "out"
    monkeypatch.setattr(sys, "argv", ["sample.py", "--checkpoint", str(ckpt), "-n", "6", "--batch-size", "4", "--steps", "4",
                                      "--seed", "7", "--prefix", str(prefix)])
    runpy.run_path(str(ROOT / "k-diffusion_b200" / "sample.py"), run_name="__main__")
    files = sorted(tmp_path.glob("out_*.png"))
    assert [f.name for f in files] == [f"out_{i:05}.png" for i in range(6)]
    # the same images through the public API: weights loaded the way sample.py:43-44 does, sample_lms as sample.py:60
    inner = K.config.make_model(K.config.load_config(ckpt)).eval().requires_grad_(False)
    inner.load_state_dict(safetorch.load_file(str(ckpt)))
    model = K.Denoiser(inner.to(DEV), sigma_data=cfg["model"]["sigma_data"])
    sigmas = S.get_sigmas_karras(4, cfg["model"]["sigma_min"], cfg["model"]["sigma_max"], rho=7., device=DEV)
    x = K.parallel.init_noise(K.parallel.sample_seeds(7, 0, 6), (3, 64, 64), cfg["model"]["sigma_max"], DEV)
    want = torch.cat([S.sample_lms(model, x[:4], sigmas, disable=True), S.sample_lms(model, x[4:6], sigmas, disable=True)])
    for i, f in enumerate(files):
        img = Image.open(f)
        assert img.size == (64, 64) and img.mode == "RGB"
        got = K.utils.from_pil_image(img)
        ref = K.utils.from_pil_image(K.utils.to_pil_image(want[i]))
        assert float((got - ref).abs().max()) <= 2.0 / 255 + 1e-6, f"image {i} differs from the API result"     # one grey level of slack


def test_driver_is_process_count_invariant_with_seed():
    """sample_images with a seed: image i depends on (seed, i) only -- emulate 1 and 2 processes on one GPU (gather = concatenation
    of what each emulated process produced, in process order, as accelerator.gather / all_gather do)."""
    cfg, sd, inner, model, z = __import__("test_gpu_parity").build("sw64")
    sigmas = S.get_sigmas_karras(3, 1e-2, 160, device=DEV)

    class Fake:
        def __init__(self, P, r, peers):
            self.num_processes, self.process_index, self.is_main_process, self.peers = P, r, r == 0, peers

        def gather(self, x):
            self.peers[self.process_index].append(x)
            return x

    single = K.evaluation.sample_images(Fake(1, 0, [[]]), model, sigmas, 8, 4, (3, 64, 64), 160.0, sampler=S.sample_heun, seed=11)
    peers = [[], []]
    for r in range(2):
        K.evaluation.sample_images(Fake(2, r, peers), model, sigmas, 8, 4, (3, 64, 64), 160.0, sampler=S.sample_heun, seed=11)
    both = torch.cat([torch.cat([peers[0][k], peers[1][k]]) for k in range(len(peers[0]))])[:8]
    assert_close(both, single, rtol=1e-4, atol=1e-5, what="1-process vs 2-process image set")
