"""The UNet with IP-Adapter cross-attention (reference ldm/modules/diffusionmodules/openaimodel_ip.py: a copy of
openaimodel.py whose only change is `from ldm.modules.attention_ip import SpatialTransformer`, :18)."""
from ldm.modules.attention_ip import SpatialTransformer
from ldm.modules.diffusionmodules import openaimodel as _base
from ldm.modules.diffusionmodules.openaimodel import (AttentionBlock, Downsample, ResBlock, TimestepBlock,  # noqa: F401
                                                      TimestepEmbedSequential, Upsample)


class UNetModel(_base.UNetModel):
    transformer_cls = SpatialTransformer
