#!/usr/bin/env python3
"""Samples from k-diffusion models -- drop-in for the reference's sample.py (same flags, same outputs: `<prefix>_<i>.png`).

    python sample.py --checkpoint model.safetensors [--config config.json] [-n 64] [--batch-size 64] [--steps 50] [--prefix out]
    python -m torch.distributed.run --nproc-per-node 8 sample.py --checkpoint ...        # one process per GPU, batch shards

Differences from the reference script (sample.py:37-66): no `accelerate` (k_diffusion.parallel.ProcessGroup provides the slice
of its interface this script uses); the denoiser and the sampler run on libkdb200 kernels; `--seed` makes the image set
independent of the number of processes; `--precision bf16` selects the tensor-core path (default fp32 = the reference's mode).
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))

import torch

import k_diffusion as K


def main():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('--batch-size', type=int, default=64, help='the batch size')
    p.add_argument('--checkpoint', type=Path, required=True, help='the checkpoint to use')
    p.add_argument('--config', type=Path, help='the model config')
    p.add_argument('-n', type=int, default=64, help='the number of images to sample')
    p.add_argument('--prefix', type=str, default='out', help='the output prefix')
    p.add_argument('--steps', type=int, default=50, help='the number of denoising steps')
    p.add_argument('--seed', type=int, default=None, help='per-sample seeds from (seed, index): images independent of the process count')
    p.add_argument('--precision', default='fp32', choices=['fp32', 'bf16'], help='token-stream arithmetic')
    p.add_argument('--sampler', default='sample_lms', help='k_diffusion.sampling function (the reference script uses sample_lms)')
    args = p.parse_args()

    config = K.config.load_config(args.config if args.config else args.checkpoint)
    model_config = config['model']
    assert len(model_config['input_size']) == 2 and model_config['input_size'][0] == model_config['input_size'][1]
    size = model_config['input_size']

    accelerator = K.parallel.ProcessGroup()
    device = accelerator.device
    print('Using device:', device, flush=True)

    import safetensors.torch as safetorch
    inner_model = K.config.make_model(config).eval().requires_grad_(False)
    inner_model.load_state_dict(safetorch.load_file(str(args.checkpoint)))
    inner_model = inner_model.to(device).set_precision(args.precision)
    accelerator.print('Parameters:', K.utils.n_params(inner_model))
    model = K.Denoiser(inner_model, sigma_data=model_config['sigma_data'])
    sigma_min, sigma_max = model_config['sigma_min'], model_config['sigma_max']

    @torch.no_grad()
    def run():
        if accelerator.is_local_main_process:
            print('Sampling...', flush=True)
        sigmas = K.sampling.get_sigmas_karras(args.steps, sigma_min, sigma_max, rho=7., device=device)
        with K.utils.eval_mode(model):
            x_0 = K.evaluation.sample_images(accelerator, model, sigmas, args.n, args.batch_size, (model_config['input_channels'], size[0], size[1]),
                                             sigma_max, sampler=getattr(K.sampling, args.sampler), seed=args.seed,
                                             disable=not accelerator.is_local_main_process)
        if accelerator.is_main_process:
            for i, out in enumerate(x_0):
                K.utils.to_pil_image(out).save(f'{args.prefix}_{i:05}.png')

    try:
        run()
    except KeyboardInterrupt:
        pass


if __name__ == '__main__':
    main()
