"""Cross-attention (77 context tokens) timing at the step's shapes. A/B: CTRLORA_ATTN_CROSS=0 = one CTA per query tile."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctrlora_b200 import ops  # noqa: E402
from tools.profile_kernels import rnd  # noqa: E402

for (B, H, nq, nk, d) in [(8, 8, 4096, 77, 40), (8, 8, 1024, 77, 80), (8, 8, 256, 77, 160), (16, 8, 4096, 77, 40), (1, 8, 4096, 77, 40)]:
    q, k, v = rnd(B * nq, H * d), rnd(B * nk, H * d), rnd(B * nk, H * d)
    nk_pad = (nk + 7) // 8 * 8
    vt = torch.zeros(B, H, d, nk_pad, device="cuda", dtype=torch.float16)
    vt[..., :nk] = v.view(B, nk, H, d).permute(0, 2, 3, 1)
    out = torch.empty(B * nq, H * d, device="cuda", dtype=torch.float16)
    for _ in range(5):
        ops.attention(q, k, vt, B, H, nq, nk, d, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        ops.attention(q, k, vt, B, H, nq, nk, d, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    mb = (2 * q.numel() * 2 + k.numel() * 4) / 1e6
    print(f"CROSS={os.environ.get('CTRLORA_ATTN_CROSS', '1')} B={B} nq={nq} nk={nk} d={d}: {us:.1f} us  ({mb / us * 1e-3:.2f} TB/s of q+out)")
