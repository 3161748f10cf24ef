"""One un-graphed finetune training step (batch 16, rank 128, SD1.5 size) inside cudaProfilerStart/Stop, for
    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
        --log-file gpurun_out/launches_train.csv python tools/profile_train.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.TRAIN_BATCH
    device = torch.device("cuda", 0)
    from ctrlora_b200 import dropin
    dropin.activate()
    from ctrlora_b200.train import FinetuneTrainer
    model = bench.build_model(device)
    trainer = FinetuneTrainer(model)
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(B, 4, 64, 64, generator=g).to(device)
    hint = torch.randn(B, 4, 64, 64, generator=g).to(device)
    ctx = torch.randn(B, 77, 768, generator=g).to(device)
    t = torch.randint(0, 1000, (B,), generator=g).to(device)
    noise = torch.randn(B, 4, 64, 64, generator=g).to(device)
    for _ in range(2):
        trainer.step(x0, hint, ctx, t, noise)
    torch.cuda.synchronize()
    if len(sys.argv) > 2:  # python tools/profile_train.py 16 gpurun_out/gemm_shapes_train.json : per-GEMM CUDA-event timings
        import json
        from ctrlora_b200 import ops
        ops._GEMM_PROFILE, ops._GEMM_SHAPES = [], []
        trainer.step(x0, hint, ctx, t, noise)
        torch.cuda.synchronize()
        recs = [dict(shape, ms=r[1].elapsed_time(r[2]), gflop=r[0] / 1e9) for shape, r in zip(ops._GEMM_SHAPES, ops._GEMM_PROFILE)]
        ops._GEMM_PROFILE = ops._GEMM_SHAPES = None
        json.dump(recs, open(sys.argv[2], "w"), indent=0)
        tot = sum(r["ms"] for r in recs)
        print(f"{len(recs)} GEMM launches, {tot:.2f} ms, {sum(r['gflop'] for r in recs) / tot:.1f} TFLOP/s average")
        sys.exit(0)
    torch.cuda.profiler.start()
    trainer.step(x0, hint, ctx, t, noise)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
