from . import flags, flops, image_transformer_v2
from .image_transformer_v2 import ImageTransformerDenoiserModelV2
