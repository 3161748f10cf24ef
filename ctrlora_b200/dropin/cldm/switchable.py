"""Drop-in for the reference's `cldm/switchable.py`: norm / conv layers that delegate to a swappable inner layer
(`norm_layer` / `conv_layer`), used by the multi-LoRA inference ControlNet.  They are parameter holders: the fused
kernels read the *effective* layer's parameters through `effective()`."""
import torch.nn as nn

from ctrlora_b200 import ops, prepare
from ctrlora_b200.prepare import effective  # noqa: F401
from ctrlora_b200.runtime import nchw_view, pixel_major
from ldm.modules.diffusionmodules.openaimodel import _Conv


class _SwitchableNorm:
    def set_norm_layer(self, norm_layer):
        self.norm_layer = norm_layer

    def copy_weights(self):
        if self.norm_layer is not None:
            self.norm_layer.weight.data.copy_(self.weight.data)
            self.norm_layer.bias.data.copy_(self.bias.data)


class SwitchableGroupNorm(_SwitchableNorm, nn.GroupNorm):
    def __init__(self, *args, norm_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm_layer = norm_layer

    def forward(self, x):
        m = effective(self)
        y = ops.groupnorm(pixel_major(x), prepare.bias_f32(m.weight), prepare.bias_f32(m.bias), m.eps, False,
                          groups=m.num_groups)
        return nchw_view(y)


class SwitchableLayerNorm(_SwitchableNorm, nn.LayerNorm):
    def __init__(self, *args, norm_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm_layer = norm_layer

    def forward(self, x):
        m = effective(self)
        return ops.layernorm(x.half().contiguous(), prepare.bias_f32(m.weight), prepare.bias_f32(m.bias), m.eps)


class SwitchableConv2d(_Conv):
    def __init__(self, *args, conv_layer=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.conv_layer = conv_layer

    def set_conv_layer(self, conv_layer):
        self.conv_layer = conv_layer

    def copy_weights(self):
        if self.conv_layer is not None:
            self.conv_layer.weight.data.copy_(self.weight.data)
            if getattr(self, "bias", None) is not None:
                self.conv_layer.bias.data.copy_(self.bias.data)
