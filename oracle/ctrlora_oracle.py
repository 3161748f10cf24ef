"""CPU ORACLE — test infrastructure, not product code.

A plain-torch fp32 restatement of the reference's denoising hot path (xyfJASON/ctrlora), written as an interpreter
over a flat state dict with the reference's parameter names.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path (ctrlora_b200) never does.

Pinned against the live reference: tools/make_golden.py imports /root/reference (through tools/ref_shims.py), runs
the unmodified reference modules on synthetic weights/inputs and commits the outputs to tests/golden/; the CPU test
tests/test_oracle_golden.py checks this file reproduces them.  The reference itself ships no tests or golden
vectors for this path, so that fixture is the pin.

Each function cites the reference file:line it restates (paths relative to /root/reference).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ primitives
def timestep_embedding(t, dim, max_period=10000):
    """ldm/modules/diffusionmodules/util.py:154-174 — [cos | sin], half = dim // 2 frequencies."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def linear(sd, p, x, lora_scale=1.0):
    """nn.Linear, or LoRACompatibleLinear when `<p>.lora_layer.*` exists: cldm/lora.py:285-291 and :70-80
    (y = W x + b + scale * up(down(x)); network_alpha is None everywhere in the reference)."""
    y = F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))
    dk = p + ".lora_layer.down.weight"
    if dk in sd:
        y = y + lora_scale * F.linear(F.linear(x, sd[dk]), sd[p + ".lora_layer.up.weight"])
    return y


def conv(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def group_norm(sd, p, x, eps):
    """GroupNorm32 (fp32 statistics), ldm/modules/diffusionmodules/util.py:202-219; 32 groups."""
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps).type(x.dtype)


def layer_norm(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


# ------------------------------------------------------------------------------------------------ blocks
def res_block(sd, p, x, emb):
    """ResBlock._forward, ldm/modules/diffusionmodules/openaimodel.py:254-274 (use_scale_shift_norm False)."""
    h = conv(sd, p + ".in_layers.2", F.silu(group_norm(sd, p + ".in_layers.0", x, 1e-5)), padding=1)
    emb_out = linear(sd, p + ".emb_layers.1", F.silu(emb))
    h = h + emb_out[:, :, None, None]
    h = conv(sd, p + ".out_layers.3", F.silu(group_norm(sd, p + ".out_layers.0", h, 1e-5)), padding=1)
    skip = conv(sd, p + ".skip_connection", x) if (p + ".skip_connection.weight") in sd else x
    return skip + h


def cross_attention(sd, p, x, context, heads):
    """CrossAttention.forward, ldm/modules/attention.py:163-194: fp32 logits, scale d_head^-0.5, softmax over keys.
    A context given as the pair [text, ip] is the IP-Adapter variant, ldm/modules/attention_ip.py:218-289: a second
    softmax over the image tokens through to_k_ip / to_v_ip, added with the module's `ip_scale` buffer before to_out
    (self-attention layers and modules without to_k_ip use the text entry only, like CrossAttention given a tensor)."""
    ip = None
    if isinstance(context, (list, tuple)):
        context, ip = context
    ctx = x if context is None else context
    q, k, v = linear(sd, p + ".to_q", x), linear(sd, p + ".to_k", ctx), linear(sd, p + ".to_v", ctx)
    b, n, c = q.shape
    d = c // heads
    split = lambda t: t.view(b, t.shape[1], heads, d).permute(0, 2, 1, 3)  # 'b n (h d) -> b h n d'
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bhid,bhjd->bhij", q.float(), k.float()) * (d ** -0.5)
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
    out = out.permute(0, 2, 1, 3).reshape(b, n, c)
    if ip is not None and (p + ".to_k_ip.weight") in sd:
        k_ip, v_ip = split(linear(sd, p + ".to_k_ip", ip.float())), split(linear(sd, p + ".to_v_ip", ip.float()))
        sim_ip = torch.einsum("bhid,bhjd->bhij", q.float(), k_ip.float()) * (d ** -0.5)
        out_ip = torch.einsum("bhij,bhjd->bhid", sim_ip.softmax(dim=-1), v_ip).permute(0, 2, 1, 3).reshape(b, n, c)
        out = out + sd[p + ".ip_scale"] * out_ip
    return linear(sd, p + ".to_out.0", out)


def feed_forward(sd, p, x):
    """FeedForward with GEGLU, ldm/modules/attention.py:49-76: proj -> (x, gate) chunk -> x * gelu(gate) -> Linear."""
    y = linear(sd, p + ".net.0.proj", x)
    a, gate = y.chunk(2, dim=-1)
    return linear(sd, p + ".net.2", a * F.gelu(gate))


def transformer_block(sd, p, x, context, heads):
    """BasicTransformerBlock._forward, ldm/modules/attention.py:271-275."""
    x = cross_attention(sd, p + ".attn1", layer_norm(sd, p + ".norm1", x), None, heads) + x
    x = cross_attention(sd, p + ".attn2", layer_norm(sd, p + ".norm2", x), context, heads) + x  # context may be [text, ip]
    x = feed_forward(sd, p + ".ff", layer_norm(sd, p + ".norm3", x)) + x
    return x


def spatial_transformer(sd, p, x, context, heads):
    """SpatialTransformer.forward (use_linear False), ldm/modules/attention.py:321-340; GroupNorm eps 1e-6 (:88-89)."""
    b, c, h, w = x.shape
    x_in = x
    y = conv(sd, p + ".proj_in", group_norm(sd, p + ".norm", x, 1e-6))
    y = y.permute(0, 2, 3, 1).reshape(b, h * w, -1)
    i = 0
    while (p + f".transformer_blocks.{i}.norm1.weight") in sd:
        y = transformer_block(sd, p + f".transformer_blocks.{i}", y, context, heads)
        i += 1
    y = y.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    return conv(sd, p + ".proj_out", y) + x_in


def _num_children(sd, p):
    idx = {int(k[len(p) + 1:].split(".")[0]) for k in sd if k.startswith(p + ".") and k[len(p) + 1:].split(".")[0].isdigit()}
    return max(idx) + 1 if idx else 0


def sequential_block(sd, p, x, emb, context, heads):
    """TimestepEmbedSequential.forward dispatch, openaimodel.py:79-87, driven by which parameter names exist."""
    for j in range(_num_children(sd, p)):
        q = f"{p}.{j}"
        if (q + ".in_layers.0.weight") in sd:
            x = res_block(sd, q, x, emb)
        elif (q + ".transformer_blocks.0.norm1.weight") in sd:
            x = spatial_transformer(sd, q, x, context, heads)
        elif (q + ".op.weight") in sd:  # Downsample, openaimodel.py:148-159: conv3x3 stride 2 pad 1
            x = conv(sd, q + ".op", x, stride=2, padding=1)
        elif (q + ".conv.weight") in sd:  # Upsample, openaimodel.py:112-118: nearest x2 then conv3x3
            x = conv(sd, q + ".conv", F.interpolate(x, scale_factor=2, mode="nearest"), padding=1)
        elif (q + ".weight") in sd and sd[q + ".weight"].dim() == 4:
            k = sd[q + ".weight"].shape[-1]
            x = conv(sd, q, x, padding=k // 2)
        else:
            continue  # parameter-free child (SiLU / Dropout / Identity)
    return x


def time_embed(sd, t, model_channels):
    """time_embed = Linear -> SiLU -> Linear, openaimodel.py:526-531 / cldm/cldm.py:131-136."""
    e = timestep_embedding(t, model_channels)
    return linear(sd, "time_embed.2", F.silu(linear(sd, "time_embed.0", e)))


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# ------------------------------------------------------------------------------------------------ networks
def controlnet_forward(sd, hint, t, context, heads, model_channels):
    """ControlNetFinetune.forward, cldm/cldm_ctrlora_finetune.py:40-54 (the 4-channel hint latent goes straight into
    input_blocks; input_hint_block is deleted at :19).  `sd` holds keys relative to control_model."""
    emb = time_embed(sd, t, model_channels)
    outs, h = [], hint.float()
    for i in range(_num_children(sd, "input_blocks")):
        h = sequential_block(sd, f"input_blocks.{i}", h, emb, context, heads)
        outs.append(conv(sd, f"zero_convs.{i}.0", h))
    h = sequential_block(sd, "middle_block", h, emb, context, heads)
    outs.append(conv(sd, "middle_block_out.0", h))
    return outs


def unet_forward(sd, x, t, context, heads, model_channels, control=None, only_mid_control=False):
    """ControlledUnetModel.forward, cldm/cldm.py:22-45.  `control` (list of 13) is consumed by pop() like the
    reference; `sd` holds keys relative to model.diffusion_model."""
    hs = []
    emb = time_embed(sd, t, model_channels)
    h = x.float()
    for i in range(_num_children(sd, "input_blocks")):
        h = sequential_block(sd, f"input_blocks.{i}", h, emb, context, heads)
        hs.append(h)
    h = sequential_block(sd, "middle_block", h, emb, context, heads)
    if control is not None:
        h = h + control.pop()
    for i in range(_num_children(sd, "output_blocks")):
        if only_mid_control or control is None:
            h = torch.cat([h, hs.pop()], dim=1)
        else:
            h = torch.cat([h, hs.pop() + control.pop()], dim=1)
        h = sequential_block(sd, f"output_blocks.{i}", h, emb, context, heads)
    h = F.silu(group_norm(sd, "out.0", h, 1e-5))
    return conv(sd, "out.2", h, padding=1)


def apply_model(sd, x_noisy, t, context, hint_latent, heads, model_channels, control_scales=None,
                only_mid_control=False):
    """ControlFinetuneLDM.apply_model AFTER the VAE stage, cldm/cldm_ctrlora_finetune.py:67-82: `hint_latent` is
    what get_first_stage_encoding(encode_first_stage(hint)) (:76-77) returns (parity boundary)."""
    unet = _sub(sd, "model.diffusion_model.")
    if hint_latent is None:
        return unet_forward(unet, x_noisy, t, context, heads, model_channels, None, only_mid_control)
    control = controlnet_forward(_sub(sd, "control_model."), hint_latent, t, context, heads, model_channels)
    scales = control_scales if control_scales is not None else [1.0] * len(control)
    control = [c * s for c, s in zip(control, scales)]
    return unet_forward(unet, x_noisy, t, context, heads, model_channels, control, only_mid_control)


# ------------------------------------------------------------------------------------------------ schedules
def make_beta_schedule_linear(n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    """ldm/modules/diffusionmodules/util.py:21-26 ('linear'): float64 linspace of sqrt, squared; numpy out."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    return betas.numpy()


def register_schedule(n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    """DDPM.register_schedule, ldm/models/diffusion/ddpm.py:138-166: fp64 numpy math, fp32 buffers."""
    betas = make_beta_schedule_linear(n_timestep, linear_start, linear_end)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return {
        "betas": f32(betas), "alphas_cumprod": f32(ac), "alphas_cumprod_prev": f32(ac_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(ac)), "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
    }


def make_ddim_timesteps(num_ddim, num_ddpm=1000):
    """util.py:46-60 ('uniform'): range(0, T, T // S) + 1."""
    c = num_ddpm // num_ddim
    return np.asarray(list(range(0, num_ddpm, c))) + 1


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta):
    """util.py:63-74. alphacums: fp32 torch tensor (cpu); returns sigmas (torch), alphas (torch), alphas_prev (numpy)."""
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def ddim_tables(schedule, S, eta=0.0):
    """DDIMSampler.make_schedule, cldm/ddim_hacked.py:23-52: the per-index scalars p_sample_ddim reads."""
    ts = make_ddim_timesteps(S, schedule["alphas_cumprod"].shape[0])
    sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(schedule["alphas_cumprod"].cpu(), ts, eta)
    return {"timesteps": ts, "sigmas": sigmas, "alphas": alphas, "alphas_prev": alphas_prev,
            "sqrt_one_minus_alphas": np.sqrt(1.0 - alphas)}


def ddim_update(x, e_t, tables, index, noise=None, temperature=1.0):
    """p_sample_ddim's update, cldm/ddim_hacked.py:208-231 (eps-parameterisation): per-step scalars become fp32
    tensors through torch.full (:208-211), so all arithmetic is fp32."""
    b = x.shape[0]
    full = lambda v: torch.full((b, 1, 1, 1), v, device=x.device)
    a_t, a_prev = full(tables["alphas"][index]), full(tables["alphas_prev"][index])
    sigma_t, sqrt_1m = full(tables["sigmas"][index]), full(tables["sqrt_one_minus_alphas"][index])
    pred_x0 = (x - sqrt_1m * e_t) / a_t.sqrt()
    dir_xt = (1.0 - a_prev - sigma_t ** 2).sqrt() * e_t
    nz = sigma_t * (noise if noise is not None else torch.zeros_like(x)) * temperature
    x_prev = a_prev.sqrt() * pred_x0 + dir_xt + nz
    return x_prev, pred_x0


def cfg_combine(e_cond, e_uncond, scale):
    """cldm/ddim_hacked.py:190-192."""
    return e_uncond + scale * (e_cond - e_uncond)


def q_sample(schedule, x0, t, noise):
    """ldm/models/diffusion/ddpm.py:356-359 with extract_into_tensor (util.py:96-99): integer gather by t."""
    a = schedule["sqrt_alphas_cumprod"].to(x0.device)[t].view(-1, 1, 1, 1)
    s = schedule["sqrt_one_minus_alphas_cumprod"].to(x0.device)[t].view(-1, 1, 1, 1)
    return a * x0 + s * noise


def p_losses(eps, noise):
    """LatentDiffusion.p_losses, ddpm.py:902-918 with logvar == 0, l_simple_weight 1, original_elbo_weight 0:
    per-sample mean over (C,H,W) of squared error, then batch mean."""
    return ((eps - noise) ** 2).mean(dim=[1, 2, 3]).mean()


def trainable_names(control_keys, zero_trainable=True, norm_trainable=True):
    """configure_optimizers' substring filter, cldm/cldm_ctrlora_finetune.py:88-100, in its if/elif order."""
    out = []
    for n in control_keys:
        if "lora_layer" in n:
            out.append(n)
        elif ("zero_convs" in n or "middle_block_out" in n) and zero_trainable:
            out.append(n)
        elif "norm" in n and norm_trainable:
            out.append(n)
    return out
