"""Minimal `LatentDiffusion` harness around the hot path: just what `ControlLDM`, `DDIMSampler` and the training
step touch (SURVEY.md §2 row 9 marks the reference's 1800-line Lightning module as the caller of the hot path, out of
scope to rewrite).  No pytorch_lightning dependency: a plain nn.Module with the attributes the samplers read.

Kept contract (reference ldm/models/diffusion/ddpm.py): schedule buffers (`betas`, `alphas_cumprod`, ... :138-192),
`q_sample` :356-359, `get_loss` :367-380, `p_losses` :885-920, `model.diffusion_model`, `first_stage_model`,
`cond_stage_model`, `scale_factor`, `parameterization`, `num_timesteps`, `device`.
The frozen VAE / CLIP stages are instantiated only if their `target:` modules are importable (they are when these
files are overlaid onto a reference checkout); otherwise they are None and the latent-space API is used directly.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from ldm.modules.diffusionmodules.util import extract_into_tensor, make_beta_schedule
from ldm.util import default, instantiate_from_config


class DiffusionWrapper(nn.Module):
    """Holds `diffusion_model` so parameter names start with `model.diffusion_model.` (reference :1312-1351)."""

    def __init__(self, diff_model_config, conditioning_key):
        super().__init__()
        self.diffusion_model = instantiate_from_config(diff_model_config)
        self.conditioning_key = conditioning_key

    def forward(self, x, t, c_concat=None, c_crossattn=None, **kw):
        if self.conditioning_key != "crossattn":
            raise NotImplementedError("only conditioning_key='crossattn' is on the CtrLoRA path")
        return self.diffusion_model(x, t, context=torch.cat(c_crossattn, 1))


class DDPM(nn.Module):
    def __init__(self, unet_config, timesteps=1000, beta_schedule="linear", loss_type="l2", monitor=None,
                 use_ema=False, first_stage_key="image", image_size=256, channels=3, log_every_t=100,
                 clip_denoised=True, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3, given_betas=None,
                 original_elbo_weight=0., v_posterior=0., l_simple_weight=1., conditioning_key=None,
                 parameterization="eps", learn_logvar=False, logvar_init=0., **ignored):
        super().__init__()
        if parameterization != "eps":
            raise NotImplementedError("only eps-parameterisation is on the CtrLoRA path")
        if use_ema:
            raise NotImplementedError("use_ema is False in every CtrLoRA config")
        self.parameterization = parameterization
        self.clip_denoised = clip_denoised
        self.log_every_t = log_every_t
        self.first_stage_key = first_stage_key
        self.image_size = image_size
        self.channels = channels
        self.use_ema = False
        self.monitor = monitor
        self.loss_type = loss_type
        self.v_posterior = v_posterior
        self.original_elbo_weight = original_elbo_weight
        self.l_simple_weight = l_simple_weight
        self.learn_logvar = learn_logvar
        self.model = DiffusionWrapper(unet_config, conditioning_key)
        self.register_schedule(given_betas, beta_schedule, timesteps, linear_start, linear_end, cosine_s)
        logvar = torch.full(fill_value=logvar_init, size=(self.num_timesteps,))
        if learn_logvar:
            self.logvar = nn.Parameter(logvar, requires_grad=True)
        else:
            self.register_buffer("logvar", logvar)

    @property
    def device(self):
        return self.betas.device

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass

    def register_schedule(self, given_betas=None, beta_schedule="linear", timesteps=1000, linear_start=1e-4,
                          linear_end=2e-2, cosine_s=8e-3):
        betas = given_betas if given_betas is not None else make_beta_schedule(
            beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end, cosine_s=cosine_s)
        alphas = 1. - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1., ac[:-1])
        self.num_timesteps = int(betas.shape[0])
        self.linear_start, self.linear_end = linear_start, linear_end
        f32 = partial(torch.tensor, dtype=torch.float32)
        for name, val in (("betas", betas), ("alphas_cumprod", ac), ("alphas_cumprod_prev", ac_prev),
                          ("sqrt_alphas_cumprod", np.sqrt(ac)), ("sqrt_one_minus_alphas_cumprod", np.sqrt(1. - ac)),
                          ("log_one_minus_alphas_cumprod", np.log(1. - ac)), ("sqrt_recip_alphas_cumprod", np.sqrt(1. / ac)),
                          ("sqrt_recipm1_alphas_cumprod", np.sqrt(1. / ac - 1))):
            self.register_buffer(name, f32(val))
        # posterior q(x_{t-1} | x_t, x_0) tables: unused by the DDIM path, kept because they are persistent buffers in
        # the reference's checkpoints (reference :168-178), i.e. part of the strict state-dict key set
        post_var = (1 - self.v_posterior) * betas * (1. - ac_prev) / (1. - ac) + self.v_posterior * betas
        self.register_buffer("posterior_variance", f32(post_var))
        self.register_buffer("posterior_log_variance_clipped", f32(np.log(np.maximum(post_var, 1e-20))))
        self.register_buffer("posterior_mean_coef1", f32(betas * np.sqrt(ac_prev) / (1. - ac)))
        self.register_buffer("posterior_mean_coef2", f32((1. - ac_prev) * np.sqrt(alphas) / (1. - ac)))

    def q_sample(self, x_start, t, noise=None):
        noise = default(noise, lambda: torch.randn_like(x_start))
        if x_start.is_cuda and x_start.dtype == torch.float32:
            from ctrlora_b200 import ops
            return ops.q_sample(x_start, noise, t, self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod)
        # host-side tensors (schedule checks on CPU): the reference's own expression
        return (extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start +
                extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    def get_loss(self, pred, target, mean=True):
        if self.loss_type == "l1":
            loss = (target - pred).abs()
        elif self.loss_type == "l2":
            loss = torch.nn.functional.mse_loss(target, pred, reduction="none")
        else:
            raise NotImplementedError(f"unknown loss type '{self.loss_type}'")
        return loss.mean() if mean else loss


class LatentDiffusion(DDPM):
    def __init__(self, first_stage_config=None, cond_stage_config=None, num_timesteps_cond=None, cond_stage_key="image",
                 cond_stage_trainable=False, concat_mode=True, cond_stage_forward=None, conditioning_key=None,
                 scale_factor=1.0, scale_by_std=False, force_null_conditioning=False, *args, **kwargs):
        self.num_timesteps_cond = default(num_timesteps_cond, 1)
        self.scale_by_std = scale_by_std
        if conditioning_key is None:
            conditioning_key = "concat" if concat_mode else "crossattn"
        super().__init__(conditioning_key=conditioning_key, *args, **kwargs)
        self.concat_mode = concat_mode
        self.cond_stage_trainable = cond_stage_trainable
        self.cond_stage_key = cond_stage_key
        self.scale_factor = scale_factor
        self.first_stage_model = self._frozen_stage(first_stage_config)
        self.cond_stage_model = self._frozen_stage(cond_stage_config)
        self.cond_stage_forward = cond_stage_forward

    @staticmethod
    def _frozen_stage(config):
        if config is None or isinstance(config, str):
            return None
        try:
            stage = instantiate_from_config(config)
        except (ImportError, ModuleNotFoundError, AttributeError, OSError) as e:  # VAE / CLIP are outside this package
            print(f"[ctrlora_b200] frozen stage {config.get('target')} not available ({type(e).__name__}): "
                  f"latent-space API only")
            return None
        stage = stage.eval()
        for p in stage.parameters():
            p.requires_grad = False
        return stage

    # ---- frozen stages (pass-through to the reference's modules when present) ------------------------------------
    def encode_first_stage(self, x):
        if self.first_stage_model is None:
            raise RuntimeError("no first_stage_model: pass 4-channel hint latents / x0 latents directly")
        return self.first_stage_model.encode(x)

    def get_first_stage_encoding(self, encoder_posterior):
        """scale_factor * posterior.sample() (reference ddpm.py get_first_stage_encoding): the drop-in posterior draws the
        noise like the reference (host RNG) and applies mean + std * noise and the scale in one kernel."""
        from ldm.modules.distributions.distributions import DiagonalGaussianDistribution
        if isinstance(encoder_posterior, DiagonalGaussianDistribution):
            return encoder_posterior.sample(scale=self.scale_factor)
        z = encoder_posterior.sample() if hasattr(encoder_posterior, "sample") else encoder_posterior
        return self.scale_factor * z

    @torch.no_grad()
    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        if self.first_stage_model is None:
            raise RuntimeError("no first_stage_model available to decode latents")
        from ldm.models.autoencoder import AutoencoderKL
        if isinstance(self.first_stage_model, AutoencoderKL):
            return self.first_stage_model.decode(z, in_scale=1. / self.scale_factor)  # folded into post_quant_conv's weights
        return self.first_stage_model.decode(1. / self.scale_factor * z)

    def get_learned_conditioning(self, c):
        if self.cond_stage_model is None:
            raise RuntimeError("no cond_stage_model: pass the [B,77,768] text context directly")
        if self.cond_stage_forward is None and hasattr(self.cond_stage_model, "encode"):
            return self.cond_stage_model.encode(c)
        return self.cond_stage_model(c)

    # ---- training objective ----------------------------------------------------------------------------------------
    def apply_model(self, x_noisy, t, cond, return_ids=False):
        cond = cond if isinstance(cond, dict) else {"c_crossattn": cond if isinstance(cond, list) else [cond]}
        return self.model(x_noisy, t, **cond)

    def forward(self, x, c, *args, **kwargs):
        t = torch.randint(0, self.num_timesteps, (x.shape[0],), device=x.device).long()
        return self.p_losses(x, c, t, *args, **kwargs)

    def p_losses(self, x_start, cond, t, noise=None):
        noise = default(noise, lambda: torch.randn_like(x_start))
        x_noisy = self.q_sample(x_start=x_start, t=t, noise=noise)
        model_output = self.apply_model(x_noisy, t, cond)
        prefix = "train" if self.training else "val"
        loss_simple = self.get_loss(model_output, noise, mean=False).mean([1, 2, 3])
        logvar_t = self.logvar[t].to(x_start.device)
        loss = (loss_simple / torch.exp(logvar_t) + logvar_t)
        loss = self.l_simple_weight * loss.mean()
        loss_dict = {f"{prefix}/loss_simple": loss_simple.mean()}
        if self.original_elbo_weight:
            raise NotImplementedError("original_elbo_weight > 0 is not on the CtrLoRA path")
        loss_dict[f"{prefix}/loss"] = loss
        return loss, loss_dict
