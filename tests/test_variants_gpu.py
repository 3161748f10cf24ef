"""Parity of the pretrain (per-task LoRA sets, switch_lora) and inference (switchable LoRA / zero-conv / norm sets,
weighted control sum) ControlNet variants against the CPU oracle, which sees the attached set through the reference's
`<linear>.lora_layer.*` key names."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu

from tolerances import TOL  # noqa: E402

KW = dict(image_size=32, in_channels=4, hint_channels=3, model_channels=32, attention_resolutions=[4, 2, 1], num_res_blocks=2,
          channel_mult=[1, 2, 4, 4], num_heads=4, use_spatial_transformer=True, transformer_depth=1, context_dim=64, legacy=False)


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def randomize_(module, seed):
    from oracle import synth
    sd = module.state_dict()
    module.load_state_dict(synth.synth_state_dict({k: tuple(v.shape) for k, v in sd.items()}, seed, "variant."))


def oracle_view(cn, attached):
    """state dict in the finetune key layout: base weights + the currently attached LoRA set (and switched-in layers)"""
    from cldm.lora import LoRACompatibleLinear
    from ctrlora_b200 import prepare
    sd = {}
    for name, mod in cn.named_modules():
        if any(tok in name for tok in ("loras_dict", "loras_list", "zero_convs_list", "norms_list")):
            continue
        eff = prepare.effective(mod)
        for pn, p in eff.named_parameters(recurse=False):
            sd[f"{name}.{pn}"] = p.detach().float().cpu()
        if isinstance(mod, LoRACompatibleLinear) and mod.lora_layer is not None:
            sd[f"{name}.lora_layer.down.weight"] = mod.lora_layer.down.weight.detach().float().cpu()
            sd[f"{name}.lora_layer.up.weight"] = mod.lora_layer.up.weight.detach().float().cpu()
    return sd


def inputs(seed=5):
    from oracle import synth
    return (synth.synth_input("hint", (2, 4, 16, 16), seed), torch.tensor([700, 3]), synth.synth_input("ctx", (2, 77, 64), seed))


def test_pretrain_switch_lora_matches_oracle():
    from ctrlora_b200 import dropin
    dropin.activate()
    from cldm.cldm_ctrlora_pretrain import ControlNetPretrain
    from oracle import ctrlora_oracle as O
    cn = ControlNetPretrain(lora_rank=8, tasks=["canny", "depth"], **KW)
    randomize_(cn, 11)
    cn = cn.cuda().eval()
    hint, t, ctx = inputs()
    outs = {}
    for task in ("canny", "depth", "canny"):  # switching back must hit the cached fold of the first set
        cn.switch_lora(task)
        with torch.no_grad():
            got = cn(hint=hint.cuda(), timesteps=t.cuda(), context=ctx.cuda())
            ref = O.controlnet_forward(oracle_view(cn, task), hint, t, ctx, 4, 32)
        errs = [rel(a, b) for a, b in zip(got, ref)]
        print(task, "max rel err %.2e" % max(errs))
        assert max(errs) < TOL["tiny_control"]
        outs.setdefault(task, got[-1].float().cpu())
    assert rel(outs["canny"], outs["depth"]) > 1e-2  # the two LoRA sets really produce different residuals


def test_inference_two_loras_weighted_sum():
    from ctrlora_b200 import dropin
    dropin.activate()
    from cldm.cldm_ctrlora_inference import ControlNetInference
    from oracle import ctrlora_oracle as O
    cn = ControlNetInference(lora_rank=8, lora_num=2, **KW)
    randomize_(cn, 12)
    cn = cn.cuda().eval()
    hint, t, ctx = inputs(6)
    for i in (0, 1):
        cn.switch_lora(i)
        with torch.no_grad():
            got = cn(hint=hint.cuda(), timesteps=t.cuda(), context=ctx.cuda())
            ref = O.controlnet_forward(oracle_view(cn, i), hint, t, ctx, 4, 32)
        errs = [rel(a, b) for a, b in zip(got, ref)]
        print("lora set", i, "max rel err %.2e" % max(errs))
        assert max(errs) < TOL["tiny_control"]
