"""Module-tree surgery shared by the CtrLoRA ControlNet variants: swap child modules in place by dotted name, in
`named_modules()` traversal order (the order `switch_lora` relies on)."""
import torch.nn as nn

from cldm.lora import LoRACompatibleLinear


def set_child(root, dotted, new):
    *path, leaf = dotted.split(".")
    parent = root
    for part in path:
        parent = parent.get_submodule(part)
    parent._modules[leaf] = new


def to_lora_linear(m, lora_layer=None):
    """LoRACompatibleLinear carrying m's weight/bias (cldm_ctrlora_finetune.py:26-32)."""
    new = LoRACompatibleLinear(m.in_features, m.out_features, bias=m.bias is not None, lora_layer=lora_layer,
                               device=m.weight.device, dtype=m.weight.dtype)
    new.weight.data.copy_(m.weight.data)
    if m.bias is not None:
        new.bias.data.copy_(m.bias.data)
    return new


def plain_linears(root, skip=()):
    """(name, module) of every nn.Linear in traversal order, excluding subtrees whose name contains a `skip` token."""
    return [(n, m) for n, m in root.named_modules()
            if isinstance(m, nn.Linear) and not any(tok in n for tok in skip)]
