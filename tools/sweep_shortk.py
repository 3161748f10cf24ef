"""Sweep block_n for the short-K (epilogue-bound) GEMMs of the step: 1x1 / linear (+residual), GEGLU, no-residual."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlora_b200 import ops  # noqa: E402
from tools.profile_kernels import rnd, timeit  # noqa: E402

SHAPES = [  # M, K, N, geglu, residual
    (32768, 320, 320, 0, 1), (32768, 320, 320, 0, 0), (8192, 640, 640, 0, 1), (2048, 1280, 1280, 0, 1),
    (32768, 1280, 320, 0, 1), (8192, 2560, 640, 0, 1), (2048, 5120, 1280, 0, 1),
    (32768, 320, 1280, 1, 0), (8192, 640, 2560, 1, 0), (2048, 1280, 5120, 1, 0),
]
if __name__ == "__main__":
    for (m, k, n, geglu, res) in SHAPES:
        a, wt = rnd(m, k), rnd(n * (2 if geglu else 1), 1, k, scale=k ** -0.5)
        bias = torch.randn(n * (2 if geglu else 1), device="cuda")
        r = rnd(m, n) if res else None
        out = torch.empty(m, n, device="cuda", dtype=torch.float16)
        fl = 2.0 * m * n * k * (2 if geglu else 1)
        byt = 2.0 * (m * k + m * n * (2 if res else 1) + wt.numel())
        row = []
        for bn in (0, 256, 192, 160, 128, 96, 64, 32):
            if geglu and bn > 128:
                continue
            try:
                ms = timeit(lambda: ops.gemm(a, wt, bias=bias, residual=r, out=out, geglu=bool(geglu), block_n=bn), n=10)
                row.append(f"{'auto' if bn == 0 else bn}:{ms * 1e3:6.1f}")
            except Exception:
                row.append(f"{bn}:  n/a")
        print(f"M={m:5d} K={k:4d} N={n:4d} geglu={geglu} res={res} [{fl / 1e9:5.1f} GF, {byt / 1e6:5.1f} MB -> {byt / 6.5e6:5.1f} us @HBM]  " + " ".join(row))
