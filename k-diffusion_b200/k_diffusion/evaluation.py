"""Data-parallel sampling driver (reference: k_diffusion/evaluation.py:80-90 `compute_features`, the loop sample.py:62 and
train.py's evaluate() run their sampler under).

The FID / KID machinery of the reference's evaluation.py (feature extractors, kernels) is out of scope; what
the sampling path needs from this module is the driver: every process samples `ceil(n / P)` items in batches, applies
`extractor_fn` (identity when the samples themselves are wanted, sample.py:62) and the per-batch results are gathered across
processes -- outside any kernel-timed region, one collective per batch, as the reference does through `accelerator.gather`.
"""
import math

import torch

from . import parallel

try:
    from tqdm.auto import trange
except ImportError:
    def trange(*args, disable=None):
        return range(*args)


def compute_features(accelerator, sample_fn, extractor_fn, n, batch_size):
    """Same contract as the reference (evaluation.py:80-90).  `accelerator` is anything with `num_processes`, `is_main_process`
    and `gather(tensor)` -- an `accelerate.Accelerator` or `k_diffusion.parallel.ProcessGroup`."""
    n_per_proc = math.ceil(n / accelerator.num_processes)
    feats_all = []
    try:
        for i in trange(0, n_per_proc, batch_size, disable=not accelerator.is_main_process):
            cur_batch_size = min(n - i, batch_size)
            samples = sample_fn(cur_batch_size)[:cur_batch_size]
            feats_all.append(accelerator.gather(extractor_fn(samples)))
    except StopIteration:
        pass
    return torch.cat(feats_all)[:n]


def sample_images(accelerator, model, sigmas, n, batch_size, shape, sigma_max, sampler=None, seed=None, extra_args_fn=None, disable=True):
    """`n` samples of `shape` = (C, H, W) with `sampler(model, x, sigmas, ...)` (default: sample_lms, what sample.py:60 calls).

    seed=None draws the initial latents from torch's global generator on the device like the reference (sample.py:59; results then
    depend on the process layout).  With a seed, latent i is a pure function of (seed, global sample index) (parallel.init_noise),
    so the image set does not depend on how many processes produced it."""
    from . import sampling
    sampler = sampling.sample_lms if sampler is None else sampler
    device = sigmas.device
    P, r = accelerator.num_processes, accelerator.process_index
    done = [0]

    def sample_fn(cur):
        if seed is None:
            x = torch.randn([cur, *shape], device=device) * sigma_max
        else:      # batch k of process r covers the global indices (k * P + r) * batch_size + [0, cur): what gather concatenates
            start = (done[0] * P + r) * batch_size
            x = parallel.init_noise(parallel.sample_seeds(seed, start, start + cur), tuple(shape), sigma_max, device)
        done[0] += 1
        extra = {} if extra_args_fn is None else extra_args_fn(cur)
        return sampler(model, x, sigmas, extra_args=extra, disable=disable)

    return compute_features(accelerator, sample_fn, lambda x: x, n, batch_size)
