"""Drop-in for the reference's `ldm/models/autoencoder.py`: `AutoencoderKL` (the frozen SD first stage) with the reference's
constructor kwargs and state-dict keys (`encoder.*`, `decoder.*`, `quant_conv.*`, `post_quant_conv.*`), inference only.

CtrLoRA calls `encode` on the condition image in EVERY apply_model (cldm/cldm_ctrlora_finetune.py:76-77) and on the target
image of every training sample; `decode` runs once per generated image.  The 1x1 `quant_conv` is composed into the encoder's
last conv in fp32 (one GEMM, the moments never round to fp16 in between); `post_quant_conv` is a GEMM whose padded
pixel-major output feeds the decoder's first conv directly."""
from contextlib import contextmanager

import torch
import torch.nn as nn

from ctrlora_b200 import ops, prepare
from ctrlora_b200.runtime import pixel_major
from ldm.modules.diffusionmodules.model import Decoder, Encoder
from ldm.modules.distributions.distributions import DiagonalGaussianDistribution
from ldm.util import instantiate_from_config


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig, lossconfig, embed_dim, ckpt_path=None, ignore_keys=[], image_key="image",
                 colorize_nlabels=None, monitor=None, ema_decay=None, learn_logvar=False):
        super().__init__()
        if ema_decay is not None or learn_logvar:
            raise NotImplementedError("VAE training (EMA / learned logvar) is outside the CtrLoRA path: the first stage is frozen")
        self.learn_logvar = learn_logvar
        self.image_key = image_key
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.loss = instantiate_from_config(lossconfig)
        assert ddconfig["double_z"]
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim
        if colorize_nlabels is not None:
            assert type(colorize_nlabels) == int
            self.register_buffer("colorize", torch.randn(3, colorize_nlabels, 1, 1))
        if monitor is not None:
            self.monitor = monitor
        self.use_ema = False
        self._prep = prepare.PrepCache()
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    @property
    def device(self):
        return self.quant_conv.weight.device

    def init_from_ckpt(self, path, ignore_keys=list()):
        sd = torch.load(path, map_location="cpu")["state_dict"]
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                print("Deleting key {} from state_dict.".format(k))
                del sd[k]
        self.load_state_dict(sd, strict=False)
        print(f"Restored from {path}")

    @contextmanager
    def ema_scope(self, context=None):
        yield None

    @torch.no_grad()
    def encode(self, x):
        """x: fp32 [B, 3, H, W] in [-1, 1] -> posterior over [B, z, H/8, W/8]  (reference :82-86)"""
        moments = self.encoder(x, post=(self.quant_conv.weight, self.quant_conv.bias))
        return DiagonalGaussianDistribution(moments)

    @torch.no_grad()
    def decode(self, z, in_scale=1.0):
        """z: fp32 [B, z, h, w] (times in_scale, so the LDM's 1/scale_factor costs nothing) -> image fp32 [B, 3, 8h, 8w]
        (reference :88-91)"""
        pq = self.post_quant_conv
        zc = z.shape[1]
        c_pad = (zc + 7) // 8 * 8

        def build():
            n = pq.out_channels
            w = torch.zeros((16, c_pad), device=pq.weight.device, dtype=torch.float32)
            w[:n, :zc] = pq.weight.detach().float().view(n, zc) * in_scale
            b = torch.zeros(16, device=pq.weight.device, dtype=torch.float32)
            b[:n] = pq.bias.detach().float()
            return ops.cast_transpose(w.contiguous(), 16 * c_pad, 1, 1).view(16, 1, c_pad), b

        w16, b16 = self._prep.get(("pq", float(in_scale)), [pq.weight, pq.bias], build)
        zp = pixel_major(z, c_pad if c_pad != zc else None)
        return self.decoder.run(ops.gemm(zp, w16, bias=b16))

    def forward(self, input, sample_posterior=True):
        posterior = self.encode(input)
        z = posterior.sample() if sample_posterior else posterior.mode()
        return self.decode(z), posterior

    def get_input(self, batch, k):
        x = batch[k]
        if len(x.shape) == 3:
            x = x[..., None]
        return x.permute(0, 3, 1, 2).to(memory_format=torch.contiguous_format).float()

    def get_last_layer(self):
        return self.decoder.conv_out.weight


class IdentityFirstStage(nn.Module):
    def __init__(self, *args, vq_interface=False, **kwargs):
        self.vq_interface = vq_interface
        super().__init__()

    def encode(self, x, *args, **kwargs):
        return x

    def decode(self, x, *args, **kwargs):
        return x

    def quantize(self, x, *args, **kwargs):
        if self.vq_interface:
            return x, None, [None, None, None]
        return x

    def forward(self, x, *args, **kwargs):
        return x
