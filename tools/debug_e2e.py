"""Host-time breakdown of bench.py's end-to-end step (H2D of the step's inputs, step, D2H of the result, host read)."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from ctrlora_b200 import dropin
dropin.activate()
from cldm.ddim_hacked import DDIMSampler
device = torch.device("cuda", 0)
model = bench.build_model(device)
sampler = DDIMSampler(model, batched_cfg=True, use_cuda_graph=True)
sampler.make_schedule(50, ddim_eta=0.0, verbose=False)
S, B = 50, bench.BATCH
gen = torch.Generator().manual_seed(100)
host = {k: torch.randn(*s, generator=gen).pin_memory() for k, s in (("x", (B, 4, 64, 64)), ("hint", (B, 4, 64, 64)), ("ctx", (B, 77, 768)), ("uc", (B, 77, 768)))}
out_host = torch.empty(B, 4, 64, 64).pin_memory(); stats_host = torch.empty(B).pin_memory()
def step(i, x, c, u):
    index = S - 1 - (i % S)
    ts = torch.full((B,), int(sampler.ddim_timesteps[index]), device=device, dtype=torch.long)
    return sampler.p_sample_ddim(x, c, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=u)
T = []
for i in range(28):
    t0 = time.perf_counter()
    d = {k: host[k].to(device, non_blocking=True) for k in host}
    c = {"c_crossattn": [d["ctx"]], "c_concat": [d["hint"]]}; u = {"c_crossattn": [d["uc"]], "c_concat": [d["hint"]]}
    t1 = time.perf_counter()
    xp, _ = step(i, d["x"], c, u)
    t2 = time.perf_counter()
    out_host.copy_(xp, non_blocking=True); stats_host.copy_(sampler.last_stats, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    t3 = time.perf_counter()
    host["x"].copy_(out_host)
    t4 = time.perf_counter()
    T.append((t1 - t0, t2 - t1, t3 - t2, t4 - t3))
for i, t in enumerate(T):
    print(i, " ".join(f"{v * 1e3:7.3f}" for v in t), f"total {sum(t) * 1e3:7.3f} ms")
import statistics
print("median h2d/enqueue/wait/hostcopy ms:", [round(statistics.median(t[j] for t in T[8:]) * 1e3, 3) for j in range(4)])
if os.environ.get("E2E_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(40):
        d = {k: host[k].to(device, non_blocking=True) for k in host}
        c = {"c_crossattn": [d["ctx"]], "c_concat": [d["hint"]]}; u = {"c_crossattn": [d["uc"]], "c_concat": [d["hint"]]}
        xp, _ = step(i, d["x"], c, u)
        out_host.copy_(xp, non_blocking=True); stats_host.copy_(sampler.last_stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        host["x"].copy_(out_host)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
