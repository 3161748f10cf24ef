"""Tensor-level wrappers over the C ABI.  torch is used for device memory and the current stream only.

Layout convention: activations are fp16 "pixel-major" tensors, either [B, H, W, C] or [M, C]; the channel dim is
contiguous.  Weights are prepared once (see ctrlora_b200.prepare) as fp16 [N, taps, Cin].
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, check


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.CtrloraError("ctrlora_b200 ops need CUDA tensors (sm_100a); there is no CPU path")


def _as_bhwc(a):
    if a.dim() == 2:
        return 1, 1, a.shape[0], a.shape[1], a.stride(0)
    assert a.dim() == 4 and a.stride(3) == 1
    b, h, w, c = a.shape
    ld = a.stride(2)
    assert a.stride(1) == w * ld and a.stride(0) == h * w * ld, "pixel-major layout required"
    return b, h, w, c, ld


def gemm(a, w, *, ksize=1, bias=None, rowbias=None, rows_per_img=0, residual=None, out_scale=1.0, a2=None, w2=None,
         geglu=False, out=None, out_f32=False, seg_outs=None, seg_width=0, transposed=(0, 0, 0), head_dim=0,
         tok_pad=0, block_n=0, simt=False):
    """out = epilogue(conv_or_linear(a, w) [+ a2 @ w2^T]); see `ctrlora_gemm_f16` in include/ctrlora_b200.h.

    a: fp16 [B,H,W,C] or [M,K]; w: fp16 [N(2N), ksize*ksize, C]; returns the output tensor ([..., N]).
    """
    _require_cuda(a, w)
    assert a.dtype == torch.float16 and w.dtype == torch.float16 and w.is_contiguous()
    b, h, wd, c, ld = _as_bhwc(a)
    n_rows = w.shape[0]
    n = n_rows // 2 if geglu else n_rows
    assert w.numel() == n_rows * ksize * ksize * c, (w.shape, ksize, c)
    M = b * h * wd
    args = GemmArgs()
    args.a, args.a_b, args.a_h, args.a_w, args.a_c, args.a_ld = _ptr(a), b, h, wd, c, ld
    args.w, args.kh, args.kw, args.pad = _ptr(w), ksize, ksize, (ksize - 1) // 2
    if a2 is not None:
        b2, h2, w2d, c2, ld2 = _as_bhwc(a2)
        assert (b2, h2, w2d) == (b, h, wd) and w2.dtype == torch.float16 and w2.is_contiguous()
        args.a2, args.a2_c, args.a2_ld, args.w2 = _ptr(a2), c2, ld2, _ptr(w2)
    args.n, args.block_n, args.geglu = n, block_n, int(geglu)
    if seg_outs is not None:
        outs = list(seg_outs)
        ldc = seg_width
        for i, o in enumerate(outs):
            args.out[i] = o.data_ptr()
            args.transposed[i] = int(transposed[i])
        ret = outs
    else:
        if out is None:
            shape = (M, n) if a.dim() == 2 else (b, h, wd, n)
            out = torch.empty(shape, device=a.device, dtype=torch.float32 if out_f32 else torch.float16)
        assert out.stride(-1) == 1
        ldc = out.stride(-2)
        args.out[0] = out.data_ptr()
        ret = out
    args.seg_width, args.ldc, args.out_f32 = seg_width, ldc, int(out_f32)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == n_rows
    if rowbias is not None:
        assert rowbias.dtype == torch.float32 and rowbias.shape[-1] == n and rowbias.is_contiguous()
    args.bias, args.rowbias, args.rows_per_img = _ptr(bias), _ptr(rowbias), rows_per_img
    if residual is not None:
        assert residual.dtype == torch.float16 and residual.stride(-1) == 1
        args.residual, args.ldr = _ptr(residual), residual.stride(-2)
    args.out_scale, args.head_dim, args.tok_pad, args.bf16 = float(out_scale), head_dim, tok_pad, 0
    lib = _lib.load()
    fn = lib.ctrlora_gemm_f16_simt if simt else lib.ctrlora_gemm_f16
    check(fn(C.byref(args), _stream()), "ctrlora_gemm_f16")
    return ret
