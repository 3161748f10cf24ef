"""Forward attention timing at the SD1.5 64x64 / 32x32 self-attention shapes (CUDA events, 20 launches after warm-up).
A/B: CTRLORA_ATTN_STREAM64=0 selects the 128-key single-S-buffer kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctrlora_b200 import ops  # noqa: E402
from tools.profile_kernels import rnd  # noqa: E402

for (B, H, nq, nk, d) in [(8, 8, 4096, 4096, 40), (16, 8, 4096, 4096, 40), (8, 8, 1024, 1024, 80)]:
    q, k, v = rnd(B * nq, H * d), rnd(B * nk, H * d), rnd(B * nk, H * d)
    vt = v.view(B, nk, H, d).permute(0, 2, 3, 1).contiguous()
    out = torch.empty(B * nq, H * d, device="cuda", dtype=torch.float16)
    lse = torch.empty(B, H, nq, device="cuda", dtype=torch.float32)
    for _ in range(5):
        ops.attention(q, k, vt, B, H, nq, nk, d, out=out, lse=lse)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ops.attention(q, k, vt, B, H, nq, nk, d, out=out, lse=lse)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    exps = B * H * nq * nk
    print(f"STREAM64={os.environ.get('CTRLORA_ATTN_STREAM64', '1')} B={B} nq={nq} d={d}: {us:.1f} us  {exps / us / 1e6:.2f} Texp/s  "
          f"{4 * exps * d / us / 1e6:.0f} TF/s")
