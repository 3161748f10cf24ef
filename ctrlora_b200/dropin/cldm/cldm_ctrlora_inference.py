"""Drop-in for the reference's `cldm/cldm_ctrlora_inference.py`: a ControlNet with `lora_num` switchable sets of
(LoRA layers, zero-convs, norm layers) and an LDM whose apply_model sums the weighted control stacks."""
import copy

import torch
import torch.nn as nn

from ctrlora_b200 import ops

from cldm.cldm import ControlLDM, ControlNet
from cldm.ddim_hacked import DDIMSampler
from cldm.lora import LoRACompatibleLinear, LoRALinearLayer
from cldm.switchable import SwitchableConv2d, SwitchableGroupNorm, SwitchableLayerNorm
from cldm._inject import plain_linears, set_child, to_lora_linear
from ctrlora_b200.runtime import Scaled, unwrap_scaled


class ControlNetInference(ControlNet):
    def __init__(self, lora_rank=128, lora_num=1, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_rank = lora_rank
        self.lora_num = lora_num
        del self.input_hint_block
        linears = plain_linears(self)
        zero_convs = [(n, m) for n, m in self.named_modules()
                      if ('zero_convs' in n or 'middle_block_out' in n) and isinstance(m, nn.Conv2d)]
        norms = [(n, m) for n, m in self.named_modules() if 'norm' in n and isinstance(m, (nn.GroupNorm, nn.LayerNorm))]
        self.loras_list = nn.ModuleList([
            nn.ModuleList([LoRALinearLayer(m.in_features, m.out_features, rank=lora_rank) for _, m in linears])
            for _ in range(lora_num)])
        self.zero_convs_list = nn.ModuleList([nn.ModuleList([copy.deepcopy(m) for _, m in zero_convs])
                                              for _ in range(lora_num)])
        self.norms_list = nn.ModuleList([nn.ModuleList([copy.deepcopy(m) for _, m in norms]) for _ in range(lora_num)])
        for n, m in linears:
            set_child(self, n, to_lora_linear(m))
        for n, m in zero_convs:
            set_child(self, n, SwitchableConv2d(m.in_channels, m.out_channels, m.kernel_size, m.stride, m.padding,
                                                m.dilation, m.groups, m.bias is not None))
        for n, m in norms:
            if isinstance(m, nn.GroupNorm):
                # note: the reference's switchable GroupNorm is built with the default eps (cldm_ctrlora_inference.py:88);
                # the swapped-in copies in norms_list keep the original eps and are the layers that get evaluated
                set_child(self, n, SwitchableGroupNorm(m.num_groups, m.num_channels))
            else:
                set_child(self, n, SwitchableLayerNorm(m.normalized_shape, m.eps, m.elementwise_affine))

    def forward(self, hint, timesteps, context, **kwargs):
        return self.forward_latent_hint(hint, timesteps, context)

    def switch_lora(self, index: int):
        lora, zero_convs, norms = self.loras_list[index], self.zero_convs_list[index], self.norms_list[index]
        i = iz = inorm = 0
        for n, m in self.named_modules():
            if isinstance(m, LoRACompatibleLinear):
                m.set_lora_layer(lora[i])
                i += 1
            elif isinstance(m, SwitchableConv2d):
                m.set_conv_layer(zero_convs[iz])
                iz += 1
            elif isinstance(m, (SwitchableGroupNorm, SwitchableLayerNorm)):
                m.set_norm_layer(norms[inorm])
                inorm += 1

    def copy_weights_to_switchable(self):
        """Push the weights loaded into the switchable shells into the currently attached inner layers (call after
        switch_lora() and load_state_dict(), as in the reference :132-139)."""
        for n, m in self.named_modules():
            if isinstance(m, (SwitchableConv2d, SwitchableGroupNorm, SwitchableLayerNorm)):
                m.copy_weights()


class ControlInferenceLDM(ControlLDM):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.lora_weights = [1.0 / self.control_model.lora_num] * self.control_model.lora_num

    @torch.no_grad()
    def sample_log(self, cond, batch_size, ddim, ddim_steps, **kwargs):
        sampler = DDIMSampler(self)
        b, c, h, w = cond["c_concat"][0].shape
        shape = (self.channels, h // 8, w // 8) if c != self.channels else (self.channels, h, w)
        return sampler.sample(ddim_steps, batch_size, shape, cond, verbose=False, **kwargs)

    @ops.with_stats_arena
    def apply_model(self, x_noisy, t, conds, *args, **kwargs):
        if isinstance(conds, dict):
            conds = [conds]
        assert isinstance(conds, (list, tuple))
        assert len(conds) == self.control_model.lora_num
        assert len(self.lora_weights) == self.control_model.lora_num
        diffusion_model = self.model.diffusion_model
        cc = conds[0]['c_crossattn']
        cond_txt = cc[0] if len(cc) == 1 else torch.cat(cc, 1)
        stacks = []
        for i, cond in enumerate(conds):
            self.control_model.switch_lora(i)
            hint = self.hint_latent(cond['c_concat'])
            stacks.append(self.control_model(hint=hint, timesteps=t, context=cond_txt))
        if len(stacks) == 1:
            control = [Scaled(c, s * self.lora_weights[0]) for c, s in zip(stacks[0], self.control_scales)]
        else:
            # sum_i w_i * scale_j * control_i[j]  (reference :172-176): one n-ary kernel per residual, fp32 accumulate
            control = [ops.weighted_sum([st[j] for st in stacks], [s * w for w in self.lora_weights])
                       for j, s in enumerate(self.control_scales)]
        return diffusion_model(x=x_noisy, timesteps=t, context=cond_txt, control=control,
                               only_mid_control=self.only_mid_control)
