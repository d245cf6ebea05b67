"""k_diffusion -- B200-native drop-in for the sampling hot path of crowsonkb/k-diffusion.

Same import surface as the reference for the path in scope (`sampling`, `layers.Denoiser`,
`external.DiscreteSchedule`, `config.load_config / make_model / make_denoiser_wrapper`,
`models.ImageTransformerDenoiserModelV2`); everything on the latent runs in libkdb200.so.
"""
from . import config, evaluation, external, layers, models, parallel, sampling, synth, utils
from .layers import Denoiser

__all__ = ["config", "evaluation", "external", "layers", "models", "parallel", "sampling", "synth", "utils", "Denoiser"]
