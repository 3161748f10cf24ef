"""cldm/cldm_style.py of the reference: cldm/cldm.py with the UNet taken from openaimodel_ip (its :16); ControlNet and
ControlLDM are the plain ones (its ControlNet is built on ldm.modules.attention, :15)."""
from cldm.cldm import ControlledForward, ControlLDM, ControlNet, _ctx16  # noqa: F401
from ctrlora_b200.runtime import to_f16_rows
from ldm.modules.diffusionmodules.openaimodel_ip import UNetModel


class ControlledUnetModel(ControlledForward, UNetModel):
    @staticmethod
    def _context(context):
        """`[[text, ip]]` (one pair per transformer depth, cldm/cldm_ctrlora_style_inference.py:184-187), a plain tensor,
        or a list of tensors: text tokens go through the registered-context cache, image tokens become fp16 rows"""
        if not isinstance(context, list):
            return _ctx16(context)
        out = []
        for c in context:
            if isinstance(c, (list, tuple)):
                txt, ip = c
                ip16 = None if ip is None else to_f16_rows(ip).view(ip.shape[0], ip.shape[1], -1)
                out.append([_ctx16(txt), ip16])
            else:
                out.append(_ctx16(c))
        return out
