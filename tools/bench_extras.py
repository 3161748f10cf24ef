"""Measurements beside bench.py's headline (one B200, CUDA events, 3 warm-ups, median of 7):
  * first-stage VAE at 512x512, batch 4: encode (posterior sample) and decode, TFLOP/s against 1116.7 / 2514.5 GF per image;
  * multi-LoRA inference (2 LoRA sets, batch 4 + CFG): ControlNet stacks grouped in one pass vs one pass per set;
  * a whole 512x512 sampling run through the public API: VAE-encode the condition image once (cache_hint_latent), 50 DDIM steps
    with CFG, VAE-decode -- seconds per batch of 4 images.
    python tools/bench_extras.py > profiles/r2_extras.json"""
import json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ctrlora_b200 import dropin
dropin.activate()
from cldm.ddim_hacked import DDIMSampler
from cldm.model import create_model

dev = torch.device("cuda", 0)


def timed(fn, n=7, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


res = {}
B = 4
model = bench.build_model(dev)  # finetune config incl. the VAE (random weights)
vae = model.first_stage_model
bench.random_weights_(vae, 3)
img = torch.tanh(torch.randn(B, 3, 512, 512, device=dev))
z = torch.randn(B, 4, 64, 64, device=dev)
with torch.no_grad():
    ms_enc = timed(lambda: model.get_first_stage_encoding(model.encode_first_stage(img)))
    ms_dec = timed(lambda: model.decode_first_stage(z))
res["vae"] = {"batch": B, "encode_ms": ms_enc, "decode_ms": ms_dec,
              "encode_tflops": B * 1116.7e9 / (ms_enc * 1e-3) / 1e12, "decode_tflops": B * 2514.5e9 / (ms_dec * 1e-3) / 1e12}

# whole sampling run from a condition IMAGE
ctx, uc = torch.randn(B, 77, 768, device=dev), torch.randn(B, 77, 768, device=dev)
model.cache_hint_latent = True
sampler = DDIMSampler(model)
cond = {"c_crossattn": [ctx], "c_concat": [img]}
ucond = {"c_crossattn": [uc], "c_concat": [img]}


def run():
    with torch.no_grad():
        lat, _ = sampler.sample(50, B, (4, 64, 64), cond, verbose=False, eta=0.0, unconditional_guidance_scale=7.5,
                                unconditional_conditioning=ucond)
        return model.decode_first_stage(lat)


ms_run = timed(run, n=3, warm=2)
res["sampling_run_512"] = {"batch": B, "ddim_steps": 50, "cfg": 7.5, "ms": ms_run, "images_per_sec": B / (ms_run * 1e-3),
                           "note": "VAE encode of the condition image once per run (cache_hint_latent), 50 steps, VAE decode"}
del sampler, model
torch.cuda.empty_cache()

# multi-LoRA: grouped vs sequential ControlNet passes
m2 = bench.build_model(dev, config=os.path.join(ROOT, "configs", "ctrlora_inference_sd15_rank128_2loras.yaml"))
x = torch.randn(2 * B, 4, 64, 64, device=dev)
t = torch.full((2 * B,), 501, device=dev, dtype=torch.long)
ctx2 = torch.randn(2 * B, 77, 768, device=dev)
conds = [{"c_crossattn": [ctx2], "c_concat": [torch.randn(2 * B, 4, 64, 64, device=dev)]} for _ in range(2)]
out = {}
for name, flag in (("grouped", True), ("sequential", False)):
    m2.grouped_multi_lora = flag
    from ctrlora_b200.graph import GraphedCallable
    fn = lambda xx, tt: m2.apply_model(xx, tt, conds)
    with torch.no_grad():
        gc = GraphedCallable(fn, [x, t])
        out[name] = timed(lambda: gc(x, t))
    del gc
res["multi_lora_2sets"] = {"batch": 2 * B, "apply_model_ms": out, "speedup": out["sequential"] / out["grouped"]}
print(json.dumps(res, indent=1))
