import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from ctrlora_b200 import dropin
dropin.activate()
from cldm.model import create_model
from cldm.ddim_hacked import DDIMSampler
from oracle import synth
GOLD = os.path.join(ROOT, "tests", "golden")
gt = torch.load(os.path.join(GOLD, "tiny_finetune_golden.pt"), weights_only=False)
rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
def build(seed_cn):
    m = create_model(os.path.join(GOLD, "tiny_finetune.yaml"), init_weights=False)
    m.control_model.load_state_dict(synth.synth_state_dict(gt["control_shapes"], seed_cn, "control_model."))
    m.model.diffusion_model.load_state_dict(synth.synth_state_dict(gt["unet_shapes"], 0, "model.diffusion_model."))
    return m.cuda().eval()
B, H = 2, 16
mk = lambda n, s: synth.synth_input(n, s, 0).cuda()
x, hint, ctx, uc = mk("x", (B, 4, H, H)), mk("hint", (B, 4, H, H)), mk("ctx", (B, 77, 64)), mk("uc_ctx", (B, 77, 64))
cond = {"c_crossattn": [ctx], "c_concat": [hint]}; ucond = {"c_crossattn": [uc], "c_concat": [hint]}
ts = torch.full((B,), 981, dtype=torch.long, device="cuda")
kw = dict(index=49, unconditional_guidance_scale=7.5, unconditional_conditioning=ucond)
def samp(model, graph):
    s = DDIMSampler(model, use_cuda_graph=graph); s.make_schedule(50, ddim_eta=0.0, verbose=False); return s
model = build(0)
se, sg = samp(model, False), samp(model, True)
xa = se.p_sample_ddim(x, cond, ts, **kw)[0].clone()
xa2 = se.p_sample_ddim(x, cond, ts, **kw)[0].clone()
xb = sg.p_sample_ddim(x, cond, ts, **kw)[0].clone()
xb2 = sg.p_sample_ddim(x, cond, ts, **kw)[0].clone()
print("old weights: eager vs eager", rel(xa2, xa), "graph vs eager", rel(xb, xa), "graph vs graph", rel(xb2, xb))
with torch.no_grad():
    ea = model.apply_model(x, ts, cond).clone(); ea2 = model.apply_model(x, ts, cond).clone()
print("apply_model eager twice", rel(ea2, ea))
model.control_model.load_state_dict(synth.synth_state_dict(gt["control_shapes"], 1, "control_model."))
xc = se.p_sample_ddim(x, cond, ts, **kw)[0].clone()
xd = sg.p_sample_ddim(x, cond, ts, **kw)[0].clone()
fresh = build(1)
xe = samp(fresh, False).p_sample_ddim(x, cond, ts, **kw)[0].clone()
xf = samp(fresh, True).p_sample_ddim(x, cond, ts, **kw)[0].clone()
print("new weights: eager(reloaded) vs fresh", rel(xc, xe), "graph(reloaded) vs fresh", rel(xd, xe), "fresh graph vs fresh eager", rel(xf, xe),
      "new vs old", rel(xe, xa))
