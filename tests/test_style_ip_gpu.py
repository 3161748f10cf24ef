"""IP-Adapter / style variant (second half): cldm.cldm_ctrlora_style_inference.ControlInferenceLDM
with the UNet of cldm.cldm_style (IPCrossAttention in every attn2), against outputs of the UNMODIFIED reference
(tests/golden/tiny_style_golden.pt, `tools/make_golden.py --style`) and against the oracle's restatement of
ldm/modules/attention_ip.py:196-289 at an SD1.5-width layer.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
pytestmark = pytest.mark.gpu

from tolerances import TOL  # noqa: E402


def rel(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-20)).item()


@pytest.fixture(scope="module")
def g():
    return torch.load(os.path.join(GOLD, "tiny_style_golden.pt"), weights_only=False)


@pytest.fixture(scope="module")
def model(g):
    from ctrlora_b200 import dropin
    dropin.activate()
    from cldm.model import create_model
    from oracle import synth
    m = create_model(os.path.join(GOLD, "tiny_style.yaml"), init_weights=False)
    m.control_model.load_state_dict(synth.synth_state_dict(g["control_shapes"], g["seed"], "control_model."), strict=True)
    unet = m.model.diffusion_model
    assert list(unet.state_dict().keys()) == g["unet_key_order"]  # to_k_ip / to_v_ip / ip_scale where the reference has them
    unet.load_state_dict(synth.synth_state_dict(g["unet_shapes"], g["seed"], "model.diffusion_model."), strict=True)
    return m.cuda().eval()


def inputs(g):
    from oracle import synth
    B, H, seed = g["B"], g["H"], g["seed"]
    mk = lambda n, s: synth.synth_input(n, s, seed).cuda()
    return dict(x=mk("x", (B, 4, H, H)), hint=mk("hint", (B, 4, H, H)), ctx=mk("ctx", (B, 77, 64)), ip=mk("ip", (B, 4, 64)),
                t=g["t"].cuda())


def set_scales(model, values):
    model.model.diffusion_model.load_state_dict({k: torch.tensor(v) for k, v in values.items()}, strict=False)


def test_style_apply_model_vs_reference(g, model):
    d = inputs(g)
    cond = [{"c_crossattn": [d["ctx"]], "c_concat": [d["hint"]], "c_ip": [d["ip"]]}]
    e = {}
    with torch.no_grad():
        set_scales(model, g["ip_scales_loaded"])
        e["loaded"] = rel(model.apply_model(d["x"], d["t"], cond), g["eps_loaded"])
        set_scales(model, {k: 0.8 for k in g["ip_scale_keys"]})
        e["all_0.8"] = rel(model.apply_model(d["x"], d["t"], cond), g["eps_all_0.8"])
        set_scales(model, g["ip_scales_some"])
        e["some"] = rel(model.apply_model(d["x"], d["t"], cond), g["eps_some"])
        e["no_ip"] = rel(model.apply_model(d["x"], d["t"], [{"c_crossattn": [d["ctx"]], "c_concat": [d["hint"]]}]), g["eps_no_ip"])
        e["no_hint"] = rel(model.apply_model(d["x"], d["t"], [{"c_crossattn": [d["ctx"]], "c_concat": [None], "c_ip": [d["ip"]]}]),
                           g["eps_no_hint"])
    print("style apply_model rel errors:", {k: "%.2e" % v for k, v in e.items()})
    assert max(e.values()) < TOL["tiny_eps"]
    # the image prompt and its per-layer scales really change the result
    assert rel(g["eps_some"], g["eps_no_ip"]) > 1e-3 and rel(g["eps_all_0.8"], g["eps_some"]) > 1e-3


def test_style_sampler_follows_ip_scale_under_cuda_graph(g, model):
    """The style app rewrites `ip_scale` per request on a live model (app/gradio_ctrlora_style_transfer.py:131-171): a
    sampler that captured its step in a CUDA graph must not replay the old scale (the buffer is part of the fingerprint)."""
    from cldm.ddim_hacked import DDIMSampler
    d = inputs(g)
    cond = {"c_crossattn": [d["ctx"]], "c_concat": [d["hint"]], "c_ip": [d["ip"]]}  # a dict, as the app passes it
    s = DDIMSampler(model, use_cuda_graph=True)
    s.make_schedule(10, ddim_eta=0.0, verbose=False)
    ts = torch.full((g["B"],), 901, dtype=torch.long, device="cuda")
    outs = {}
    with torch.no_grad():
        for name, scales in (("some", g["ip_scales_some"]), ("all", {k: 0.8 for k in g["ip_scale_keys"]}), ("some2", g["ip_scales_some"])):
            set_scales(model, scales)
            for _ in range(2):  # second call replays the graph
                x_prev, _ = s.p_sample_ddim(d["x"], cond, ts, index=9)
            outs[name] = x_prev.clone()
            assert s._graph is not None, "the step did not go through a CUDA graph"
    assert rel(outs["some"], outs["all"]) > 1e-4, "graph replayed a stale ip_scale"
    assert torch.equal(outs["some"], outs["some2"])
    # CFG with an image prompt on both branches (the app's cond / un_cond pair): batched into one pass
    uc = {"c_crossattn": [d["ctx"].flip(0).contiguous()], "c_concat": [d["hint"]], "c_ip": [torch.zeros_like(d["ip"])]}
    sb = DDIMSampler(model, batched_cfg=True, use_cuda_graph=True)
    sb.make_schedule(10, ddim_eta=0.0, verbose=False)
    ss = DDIMSampler(model, batched_cfg=False, use_cuda_graph=False)
    ss.make_schedule(10, ddim_eta=0.0, verbose=False)
    with torch.no_grad():
        a, _ = sb.p_sample_ddim(d["x"], cond, ts, index=9, unconditional_guidance_scale=5.0, unconditional_conditioning=uc)
        b, _ = ss.p_sample_ddim(d["x"], cond, ts, index=9, unconditional_guidance_scale=5.0, unconditional_conditioning=uc)
    assert rel(a, b) < 2e-3


@pytest.mark.parametrize("B,N,dim,heads,ctx_dim,n_ip,scale", [(2, 4096, 320, 8, 768, 4, 1.0), (3, 256, 1280, 8, 768, 16, 0.6),
                                                             (1, 1024, 640, 8, 768, 4, 0.0)])
def test_ip_cross_attention_layer_vs_oracle(B, N, dim, heads, ctx_dim, n_ip, scale):
    """one IPCrossAttention at SD1.5 widths (d_head 40 / 160 / 80), 77 text + 4 or 16 image tokens, vs the oracle"""
    from ctrlora_b200 import dropin
    dropin.activate()
    from ldm.modules.attention_ip import IPCrossAttention
    from oracle import ctrlora_oracle as O
    torch.manual_seed(dim + n_ip)
    m = IPCrossAttention(dim, context_dim=ctx_dim, heads=heads, dim_head=dim // heads)
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(2.0)
        m.ip_scale.fill_(scale)
    sd = {"a." + k: v.clone() for k, v in m.state_dict().items()}
    x, txt, ip = torch.randn(B, N, dim), torch.randn(B, 77, ctx_dim), torch.randn(B, n_ip, ctx_dim)
    m = m.cuda()
    with torch.no_grad():
        ref = O.cross_attention(sd, "a", x, [txt, ip], heads)
        got = m(x.cuda(), context=[txt.cuda(), ip.cuda()])
        ref_plain = O.cross_attention(sd, "a", x, txt, heads)
        got_plain = m(x.cuda(), context=txt.cuda())
    e, e0 = rel(got, ref), rel(got_plain, ref_plain)
    print(f"IPCrossAttention dim {dim} N {N} ip tokens {n_ip} scale {scale}: {e:.2e} (text only {e0:.2e})")
    assert e < 2e-3 and e0 < 2e-3
    if scale:
        assert rel(ref, ref_plain) > 1e-2
