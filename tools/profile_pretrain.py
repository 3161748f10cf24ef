"""One un-graphed pretraining step (batch 8, 9 tasks, rank 128, SD1.5 size) inside cudaProfilerStart/Stop, for
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches_pretrain.csv python tools/profile_pretrain.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.PRETRAIN_BATCH
    device = torch.device("cuda", 0)
    from ctrlora_b200 import dropin
    dropin.activate()
    from ctrlora_b200.train import PretrainTrainer
    model = bench.build_model(device, config=os.path.join(ROOT, "configs", "ctrlora_pretrain_sd15_9tasks_rank128.yaml"))
    trainer = PretrainTrainer(model)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(B, 4, 64, 64, generator=g).to(device)
    hint = torch.randn(B, 4, 64, 64, generator=g).to(device)
    ctx = torch.randn(B, 77, 768, generator=g).to(device)
    t = torch.randint(0, 1000, (B,), generator=g).to(device)
    noise = torch.randn(B, 4, 64, 64, generator=g).to(device)
    for _ in range(2):
        trainer.step(x0, hint, ctx, t, noise, task="canny")
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    trainer.step(x0, hint, ctx, t, noise, task="canny")
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
