"""Glue shared by the drop-in modules: the internal activation convention and the small carrier objects that let
fused kernels see across module boundaries without changing the reference's call signatures.

Internal activations are fp16 "pixel-major" buffers [B, H, W, C]; between modules they travel as the zero-copy
logical-NCHW view of that buffer (`buf.permute(0, 3, 1, 2)`, i.e. a channels_last tensor), so every module still
receives and returns [B, C, H, W]-shaped tensors like the reference's.
"""
import torch

from . import ops


def pixel_major(x, c_pad=None):
    """[B,C,H,W] tensor (fp16 channels_last view, or fp32 NCHW from outside) -> fp16 [B,H,W,C] contiguous buffer."""
    if x.dim() != 4:
        raise ValueError(f"expected a [B,C,H,W] tensor, got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("ctrlora_b200 runs on CUDA (sm_100a) only: move the model and inputs to the GPU")
    if x.dtype == torch.float16:
        v = x.permute(0, 2, 3, 1)
        if v.is_contiguous() and c_pad in (None, x.shape[1]):
            return v
        x = x.float()
    return ops.nchw_to_nhwc_f16(x.float().contiguous(), c_pad)


def nchw_view(buf):
    """fp16 [B,H,W,C] buffer -> logical [B,C,H,W] view (no copy)."""
    return buf.permute(0, 3, 1, 2)


def to_f16_rows(t):
    """fp32/fp16 [..., K] -> fp16 [rows, K] contiguous (context tokens)."""
    k = t.shape[-1]
    if t.dtype == torch.float16:
        return t.reshape(-1, k).contiguous()
    t = t.float().contiguous()
    return ops.cast_transpose(t, t.numel(), 1, 1).view(-1, k)


class Scaled:
    """A ControlNet residual with its control_scale (cldm/cldm_ctrlora_finetune.py:79): the multiply is applied where
    the residual is consumed (the GroupNorm kernel's addend scale) instead of as a separate pass."""
    __slots__ = ("tensor", "scale")

    def __init__(self, tensor, scale=1.0):
        self.tensor, self.scale = tensor, float(scale)


def unwrap_scaled(c):
    return (c.tensor, c.scale) if isinstance(c, Scaled) else (c, 1.0)


class CatSpec:
    """Deferred `cat([h (+ s1*add1), skip (+ s2*add2)], dim=1)` (cldm/cldm.py:34-42): consumed by the next ResBlock's
    GroupNorm kernel, which reads the pieces in place."""
    __slots__ = ("x1", "add1", "s1", "x2", "add2", "s2")

    def __init__(self, x1, add1=None, s1=1.0, x2=None, add2=None, s2=1.0):
        self.x1, self.add1, self.s1, self.x2, self.add2, self.s2 = x1, add1, s1, x2, add2, s2


class EmbPack:
    """Time embedding plus the per-ResBlock `emb_layers` outputs of a whole network, produced by one batched GEMV."""
    __slots__ = ("raw", "slices")

    def __init__(self, raw, slices=None):
        self.raw, self.slices = raw, slices or {}
