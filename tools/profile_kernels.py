"""Run the hot kernels at the bench workload's real shapes (batch 8 = batch 4 x cond/uncond), for `ncu --set full`.

    ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 6 -o gpurun_out/prof_gemm \
        python tools/profile_kernels.py gemm
    ncu --set full --clock-control none --import-source on -k regex:attention_kernel -c 2 -o gpurun_out/prof_attn \
        python tools/profile_kernels.py attn
Without ncu it prints CUDA-event timings (3 warm-ups, 10 timed launches each, L2 flushed between launches).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlora_b200 import ops  # noqa: E402

FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    FLUSH.zero_()


def timeit(fn, n=10):
    if os.environ.get("CTRLORA_PROFILE_ONCE") == "1":  # under ncu: one warm launch + one measured-by-ncu launch
        fn()
        flush_l2()
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device="cuda") * scale).half()


def gemm_cases():
    B = 8
    cases = [  # (name, B, H, W, C, N, ksize, residual, geglu)
        ("1x1 64x64 320->320 +res", B, 64, 64, 320, 320, 1, True, False),
        ("3x3 64x64 320->320 +res", B, 64, 64, 320, 320, 3, True, False),
        ("3x3 8x8 1280->1280 +res", B, 8, 8, 1280, 1280, 3, True, False),
        ("geglu 64x64 320->1280", B, 64, 64, 320, 1280, 1, False, True),
        ("1x1 32x32 640->640 +res", B, 32, 32, 640, 640, 1, True, False),
        ("1x1 16x16 1280->1280 +res", B, 16, 16, 1280, 1280, 1, True, False),
        ("3x3 32x32 640->640 +res", B, 32, 32, 640, 640, 3, True, False),
        ("3x3 16x16 1280->1280 +res", B, 16, 16, 1280, 1280, 3, True, False),
    ]
    for name, b, h, w, c, n, ks, res, geglu in cases:
        a = rnd(b, h, w, c)
        wt = rnd(n * (2 if geglu else 1), ks * ks, c, scale=(ks * ks * c) ** -0.5)
        bias = torch.randn(n * (2 if geglu else 1), device="cuda")
        r = rnd(b * h * w, n) if res else None
        out = torch.empty(b, h, w, n, device="cuda", dtype=torch.float16)
        fn = lambda: ops.gemm(a, wt, ksize=ks, bias=bias, residual=r, geglu=geglu, out=out)
        ms = timeit(fn)
        fl = 2.0 * b * h * w * wt.shape[0] * c * ks * ks
        byts = (a.numel() + wt.numel() + out.numel() + (r.numel() if res else 0)) * 2
        print(f"gemm {name:32s} {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  {byts / ms / 1e6:7.1f} GB/s(alg)")


def attn_cases():
    B, H = 8, 8
    for nq, nk, d in [(4096, 4096, 40), (1024, 1024, 80), (256, 256, 160), (4096, 77, 40)]:
        q, k = rnd(B * nq, H * d), rnd(B * nk, H * d)
        nk_pad = (nk + 7) // 8 * 8
        vt = rnd(B, H, d, nk_pad)
        out = torch.empty(B * nq, H * d, device="cuda", dtype=torch.float16)
        fn = lambda: ops.attention(q, k, vt, B, H, nq, nk, d, out=out)
        ms = timeit(fn)
        fl = 4.0 * B * H * nq * nk * d
        print(f"attn nq={nq} nk={nk} d={d}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  "
              f"{B * H * nq * nk / ms / 1e6:7.1f} Gexp/s")


def attn_bwd_cases():
    H = 8
    for B, nq, nk, d in [(16, 4096, 4096, 40), (16, 1024, 1024, 80), (16, 4096, 77, 40)]:
        q, k, v = rnd(B * nq, H * d), rnd(B * nk, H * d), rnd(B * nk, H * d)
        o, do = rnd(B * nq, H * d), rnd(B * nq, H * d)
        lse = torch.randn(B, H, nq, device="cuda") + 8.0
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        ms = timeit(lambda: ops.attention_bwd(q, k, v, o, do, lse, B, H, nq, nk, d, dq=dq, dk=dk, dv=dv))
        fl = 10.0 * B * H * nq * nk * d
        print(f"attn_bwd B={B} nq={nq} nk={nk} d={d}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s  "
              f"{2 * B * H * nq * nk / ms / 1e6:7.1f} Gexp/s")


def norm_cases():
    B = 8
    for h, c in [(64, 320), (32, 640), (16, 1280), (64, 640), (8, 1280)]:
        x = rnd(B, h, h, c)
        g, b = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
        ms = timeit(lambda: ops.groupnorm(x, g, b, 1e-5, True))
        print(f"groupnorm {h}x{h}x{c}: {ms * 1e3:7.1f} us  {x.numel() * 4 / ms / 1e6:7.1f} GB/s (1R+1W)")
    for m, c in [(32768, 320), (8192, 640), (2048, 1280)]:
        x = rnd(m, c)
        g, b = torch.randn(c, device="cuda"), torch.randn(c, device="cuda")
        ms = timeit(lambda: ops.layernorm(x, g, b))
        print(f"layernorm {m}x{c}: {ms * 1e3:7.1f} us  {x.numel() * 4 / ms / 1e6:7.1f} GB/s (1R+1W)")


def ddim_cases():
    for B in (4, 64):
        x, ec, eu = (torch.randn(B, 4, 64, 64, device="cuda") for _ in range(3))
        stats = torch.empty(B, device="cuda")
        ms = timeit(lambda: ops.ddim_update(x, ec, eu, 7.5, 0.5, 0.6, 0.0, 0.7, stats=stats))
        print(f"ddim_update B={B}: {ms * 1e3:7.1f} us  {x.numel() * 4 * 5 / ms / 1e6:7.1f} GB/s (3R+2W = 327 680 B/img)")


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "attn", "norm"]
    if "ddim" in which:
        ddim_cases()
    if "gemm" in which:
        gemm_cases()
    if "attn" in which:
        attn_cases()
    if "attnbwd" in which:
        attn_bwd_cases()
    if "norm" in which:
        norm_cases()
