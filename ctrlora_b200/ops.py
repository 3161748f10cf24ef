"""Tensor-level wrappers over the C ABI.  torch is used for device memory and the current stream only.

Layout convention: activations are fp16 "pixel-major" tensors, either [B, H, W, C] or [M, C]; the channel dim is
contiguous.  Weights are prepared once (see ctrlora_b200.prepare) as fp16 [N, taps, Cin].
"""
import ctypes as C

import torch

from . import _lib
from ._lib import GemmArgs, check


LAUNCHES = 0      # kernels launched through this module since the last reset (bench.py's gpu_launches)
_GEMM_PROFILE = None  # list of (flops, start_event, end_event) while profile_gemm() is active
_GEMM_SHAPES = None   # optional parallel list of problem shapes (tools/profile_step.py)
_GEMM_RECORD = None   # list of (args struct, flops, keep-alive tensors) while replay_gemms() records a step


def _count(n=1):
    global LAUNCHES
    LAUNCHES += n


def _ungraphed(fn, sampler):
    old = getattr(sampler, "use_cuda_graph", False)
    if sampler is not None:
        sampler.use_cuda_graph = False
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        if sampler is not None:
            sampler.use_cuda_graph = old


def count_launches(fn, sampler=None):
    """Kernel launches of one call of `fn` (run without CUDA-graph replay so every C-ABI call is seen)."""
    global LAUNCHES
    LAUNCHES = 0
    _ungraphed(fn, sampler)
    return LAUNCHES


def profile_gemm(fn, sampler=None):
    """Per-launch CUDA-event timing of every ctrlora_gemm_f16 launch inside `fn`: algorithmic flops and device ms."""
    global _GEMM_PROFILE
    _GEMM_PROFILE = []
    try:
        _ungraphed(fn, sampler)
        recs = _GEMM_PROFILE
    finally:
        _GEMM_PROFILE = None
    return {"launches": len(recs), "flops": float(sum(r[0] for r in recs)),
            "ms": float(sum(r[1].elapsed_time(r[2]) for r in recs))}


def replay_gemms(fn, sampler=None, reps=5):
    """Device time of the GEMM launches of `fn` replayed back to back from one CUDA graph.

    `fn` runs once un-graphed while every ctrlora_gemm_f16 argument block is recorded (its tensors are kept alive), then
    exactly those launches are captured into a graph and replayed `reps` times between two CUDA events on the capture
    stream.  Unlike per-launch events on an un-graphed step (host-bound: the GPU idles between launches, clocks and L2
    state differ from the real run) this times the kernels under the conditions they run in: PDL-chained, warm clocks.
    Returns {"launches", "flops", "ms"} for ONE pass over the step's GEMMs."""
    global _GEMM_RECORD
    _GEMM_RECORD = []
    try:
        _ungraphed(fn, sampler)
        recs = _GEMM_RECORD
    finally:
        _GEMM_RECORD = None
    lib = _lib.load()

    def launch_all():
        sp = _sp()
        for args, _, _ in recs:
            check(lib.ctrlora_gemm_f16(C.addressof(args), sp), "ctrlora_gemm_f16 (replay)")

    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        launch_all()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            launch_all()
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
    torch.cuda.current_stream().wait_stream(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    out = {"launches": len(recs), "flops": float(sum(r[1] for r in recs)), "ms": float(ms)}
    del graph, recs
    return out


_SPLITK = {}  # device index -> (fp32 workspace, uint32 counters); zero on entry and on exit of every GEMM launch
SPLITK_WS_BYTES = 128 << 20  # one fp32 [1280, 9*1280] dense-conv gradient slice is 59 MB
SPLITK_COUNTERS = 4096


def _splitk_buffers(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SPLITK:
        _SPLITK[key] = (torch.zeros(SPLITK_WS_BYTES // 4, device=device, dtype=torch.float32),
                        torch.zeros(SPLITK_COUNTERS, device=device, dtype=torch.int32))
    return _SPLITK[key]


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


def gemm(a, w, *, ksize=1, bias=None, rowbias=None, rows_per_img=0, rowbias_ld=0, residual=None, out_scale=1.0, a2=None,
         w2=None,
         geglu=False, out=None, out_f32=False, seg_outs=None, seg_width=0, transposed=(0, 0, 0), head_dim=0,
         tok_pad=0, block_n=0, split_k=0, dup_out=None, single_cta=False, simt=False):
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
        assert rowbias.dtype == torch.float32 and rowbias.stride(-1) == 1
        rowbias_ld = rowbias_ld or rowbias.stride(0)
    args.bias, args.rowbias, args.rows_per_img, args.rowbias_ld = _ptr(bias), _ptr(rowbias), rows_per_img, rowbias_ld
    if residual is not None:
        assert residual.dtype in (torch.float16, torch.float32) and residual.stride(-1) == 1
        args.residual, args.ldr = _ptr(residual), residual.stride(-2)
        args.residual_f32 = int(residual.dtype == torch.float32)
    args.out_scale, args.head_dim, args.tok_pad, args.bf16 = float(out_scale), head_dim, tok_pad, 0
    ws, cnt = _splitk_buffers(a.device)
    args.split_k, args.splitk_ws, args.splitk_ws_bytes = split_k, ws.data_ptr(), SPLITK_WS_BYTES
    args.splitk_counters, args.splitk_counters_len = cnt.data_ptr(), SPLITK_COUNTERS
    if dup_out is not None:
        args.dup_out, args.dup_ld = dup_out.data_ptr(), dup_out.stride(0)
    args.force_single_cta = int(single_cta)
    lib = _lib.load()
    fn = lib.ctrlora_gemm_f16_simt if simt else lib.ctrlora_gemm_f16
    _count()
    if _GEMM_RECORD is not None and not simt:
        ktot = ksize * ksize * c + (args.a2_c if a2 is not None else 0)
        _GEMM_RECORD.append((args, 2.0 * M * n_rows * ktot,
                             (a, w, a2, w2, bias, rowbias, residual, out, seg_outs, dup_out, ws, cnt)))
    if _GEMM_PROFILE is not None and not simt:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(fn(C.addressof(args), _sp()), "ctrlora_gemm_f16")
        e1.record()
        ktot = ksize * ksize * c + (args.a2_c if a2 is not None else 0)
        _GEMM_PROFILE.append((2.0 * M * n_rows * ktot, e0, e1))
        if _GEMM_SHAPES is not None:
            _GEMM_SHAPES.append({"B": b, "H": h, "W": wd, "C": c, "N": n_rows, "ksize": ksize, "geglu": int(geglu),
                                 "c2": int(args.a2_c) if a2 is not None else 0, "segs": seg_width,
                                 "res": int(residual is not None)})
        return ret
    check(fn(C.addressof(args), _sp()), "ctrlora_gemm_f16")
    return ret


def _dp(t):
    """raw device pointer (or NULL) for argtypes-declared entry points"""
    return t.data_ptr() if t is not None else None


def _sp():
    return torch.cuda.current_stream().cuda_stream


# ---- GroupNorm statistics arena: ONE memset per step instead of one per GroupNorm ---------------------------------------
# `stats_arena_begin()` (called by the sampler / trainer at the top of a step, also inside graph capture) zeroes a flat
# fp32 buffer; every groupnorm / groupnorm_bwd call of the step then takes its {sum, sumsq} workspace from it and tells the
# library the workspace is already zero, which removes the memset node in front of each statistics kernel and lets that
# kernel chain programmatically (PDL) onto its predecessor.  Outside a step (tests, ad-hoc calls) nothing changes.
_ARENA = {}  # device index -> [tensor, offset, active]
ARENA_FLOATS = 1 << 20


def stats_arena_begin(device=None):
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ent = _ARENA.get(key)
    if ent is None:
        if torch.cuda.is_current_stream_capturing():
            return  # never allocate the arena inside a capture (it would live in that graph's pool); plain path instead
        ent = _ARENA[key] = [torch.zeros(ARENA_FLOATS, device=dev, dtype=torch.float32), 0, True]
    ent[0].zero_()
    ent[1], ent[2] = 0, True
    _count()


def stats_arena_end(device=None):
    key = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
    if key in _ARENA:
        _ARENA[key][2] = False


def with_stats_arena(fn):
    """Decorator for a step-level entry point (apply_model): GroupNorm statistics come from the per-step arena."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, x, *args, **kwargs):
        nested = False
        if torch.is_tensor(x) and x.is_cuda:
            key = x.device.index if x.device.index is not None else torch.cuda.current_device()
            nested = key in _ARENA and _ARENA[key][2]
            if not nested:
                stats_arena_begin(x.device)
        try:
            return fn(self, x, *args, **kwargs)
        finally:
            if torch.is_tensor(x) and x.is_cuda and not nested:
                stats_arena_end(x.device)
    return wrapped


def _stats_take(device, n):
    """(workspace, prezeroed) for n floats"""
    key = device.index if device.index is not None else torch.cuda.current_device()
    ent = _ARENA.get(key)
    n_al = (n + 31) // 32 * 32
    if ent is not None and ent[2] and ent[1] + n_al <= ARENA_FLOATS:
        ws = ent[0][ent[1]:ent[1] + n]
        ent[1] += n_al
        return ws, 1
    return torch.empty(n, device=device, dtype=torch.float32), 0


_GN_PARTIAL = {}  # device index -> (fp32 scratch for per-block partial statistics, uint32 self-cleaning arrival counters)
GN_PARTIAL_FLOATS = 1 << 19
GN_PARTIAL_COUNTERS = 256


def _gn_partial_buffers(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _GN_PARTIAL:
        _GN_PARTIAL[key] = (torch.empty(GN_PARTIAL_FLOATS, device=device, dtype=torch.float32),
                            torch.zeros(GN_PARTIAL_COUNTERS, device=device, dtype=torch.int32))
    return _GN_PARTIAL[key]


def groupnorm(x1, gamma, beta, eps, silu, *, add1=None, add1_scale=1.0, x2=None, add2=None, add2_scale=1.0,
              groups=32, want_raw=False, stats_ws=None, want_stats=False, out=None):
    """GroupNorm(+SiLU) over [x1 (+s1*add1) | x2 (+s2*add2)], pixel-major fp16 [B,H,W,C*]; returns y (and raw concat)."""
    _require_cuda(x1, x2, add1, add2)
    b, h, w, c1, ld1 = _as_bhwc(x1)
    c2, ld2 = 0, 0
    if x2 is not None:
        b2, h2, w2, c2, ld2 = _as_bhwc(x2)
        assert (b2, h2, w2) == (b, h, w)
    for ad, ref in ((add1, x1), (add2, x2)):
        if ad is not None:
            assert ad.shape == ref.shape and ad.stride() == ref.stride() and ad.dtype == torch.float16
    ctot = c1 + c2
    y = torch.empty((b, h, w, ctot), device=x1.device, dtype=torch.float16) if out is None else out
    assert y.shape == (b, h, w, ctot) and y.is_contiguous() and y.dtype == torch.float16
    raw = torch.empty_like(y) if want_raw else None
    prezeroed = 0
    if stats_ws is None:
        stats_ws, prezeroed = _stats_take(x1.device, b * groups * 2)
    a = _lib.GroupNormArgs()
    a.stats_prezeroed = prezeroed
    a.x1, a.add1, a.add1_scale, a.c1, a.ld1 = _dp(x1), _dp(add1), float(add1_scale), c1, ld1
    a.x2, a.add2, a.add2_scale, a.c2, a.ld2 = _dp(x2), _dp(add2), float(add2_scale), c2, ld2
    a.batch, a.hw, a.groups = b, h * w, groups
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.numel() == ctot
    a.gamma, a.beta, a.eps, a.silu = _dp(gamma), _dp(beta), float(eps), int(silu)
    a.y, a.raw_out, a.stats_ws = _dp(y), _dp(raw), _dp(stats_ws)
    if not torch.cuda.is_current_stream_capturing() or x1.device.index in _GN_PARTIAL or torch.cuda.current_device() in _GN_PARTIAL:
        pws, pcnt = _gn_partial_buffers(x1.device)  # (never first allocated inside a capture)
        a.partial_ws, a.partial_ws_floats = _dp(pws), GN_PARTIAL_FLOATS
        a.partial_counters, a.partial_counters_len = _dp(pcnt), GN_PARTIAL_COUNTERS
    _count(2)
    check(_lib.load().ctrlora_groupnorm_f16(C.addressof(a), _sp()), "ctrlora_groupnorm_f16")
    if want_stats:
        return (y, raw, stats_ws) if want_raw else (y, stats_ws)
    return (y, raw) if want_raw else y


def zeros(shape, device, dtype=torch.float16):
    """torch.empty + cudaMemsetAsync (a memset node; torch.zeros launches an ATen fill kernel)"""
    t = torch.empty(shape, device=device, dtype=dtype)
    check(_lib.load().ctrlora_memset_zero(_dp(t), t.numel() * t.element_size(), _sp()), "memset_zero")
    return t


def layernorm(x, gamma, beta, eps=1e-5):
    """x fp16 [..., C] with contiguous last dim and uniform row stride."""
    _require_cuda(x)
    cols = x.shape[-1]
    x2 = x.reshape(-1, cols)
    y = torch.empty((x2.shape[0], cols), device=x.device, dtype=torch.float16)
    _count(1)
    check(_lib.load().ctrlora_layernorm_f16(_dp(x2), x2.stride(0), _dp(y), cols, x2.shape[0], cols, _dp(gamma), _dp(beta),
                                            float(eps), _sp()), "ctrlora_layernorm_f16")
    return y.view(x.shape)


def attention(q, k, vt, batch, heads, nq, nk, head_dim, out=None, lse=None):
    """q [batch*nq, heads*d], k [batch*nk, heads*d], vt [batch, heads, d, nk_pad] (fp16) -> [batch*nq, heads*d]."""
    _require_cuda(q, k, vt)
    if out is None:
        out = torch.empty((batch * nq, heads * head_dim), device=q.device, dtype=torch.float16)
    _count(1)
    check(_lib.load().ctrlora_attention_f16(_dp(q), q.stride(0), _dp(k), k.stride(0), _dp(vt), vt.shape[-1], _dp(out),
                                            out.stride(0), _dp(lse), batch, heads, nq, nk, head_dim, _sp()),
          "ctrlora_attention_f16")
    return out


def nchw_to_nhwc_f16(x, c_pad=None, out=None):
    """fp32 [B,C,H,W] contiguous -> fp16 [B,H,W,c_pad] (zero-padded channels)."""
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    b, c, h, w = x.shape
    c_pad = c_pad or c
    y = torch.empty((b, h, w, c_pad), device=x.device, dtype=torch.float16) if out is None else out
    assert y.shape == (b, h, w, c_pad) and y.is_contiguous() and y.dtype == torch.float16
    _count(1)
    check(_lib.load().ctrlora_nchw_f32_to_nhwc_f16(_dp(x), _dp(y), b, c, h * w, c_pad, _sp()), "nchw_to_nhwc")
    return y


def nhwc_to_nchw_f32(x, channels=None):
    """[B,H,W,ld] fp16/fp32 pixel-major -> fp32 [B,channels,H,W]."""
    _require_cuda(x)
    b, h, w, c, ld = _as_bhwc(x)
    channels = channels or c
    y = torch.empty((b, channels, h, w), device=x.device, dtype=torch.float32)
    _count(1)
    check(_lib.load().ctrlora_nhwc_to_nchw_f32(_dp(x), int(x.dtype == torch.float32), ld, _dp(y), b, channels, h * w,
                                               _sp()), "nhwc_to_nchw")
    return y


def timestep_embedding(t, freqs):
    """t int64 [B] (device), freqs fp32 [half] (device) -> fp32 [B, 2*half]"""
    _require_cuda(t, freqs)
    assert t.dtype == torch.int64 and freqs.dtype == torch.float32
    out = torch.empty((t.shape[0], 2 * freqs.shape[0]), device=t.device, dtype=torch.float32)
    _count(1)
    check(_lib.load().ctrlora_timestep_embedding(_dp(t), _dp(freqs), _dp(out), t.shape[0], freqs.shape[0], _sp()),
          "timestep_embedding")
    return out


def small_linear(x, w, bias, silu_in=False, silu_out=False, out=None):
    """x fp32 [rows, K] (row stride free), w fp16 [N, K] -> fp32 [rows, N]"""
    _require_cuda(x, w)
    assert x.dtype == torch.float32 and w.dtype == torch.float16 and w.is_contiguous() and x.stride(1) == 1
    rows, k = x.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((rows, n), device=x.device, dtype=torch.float32)
    _count(1)
    check(_lib.load().ctrlora_small_linear(_dp(x), x.stride(0), _dp(w), _dp(bias), _dp(out), out.stride(0), rows, n, k,
                                           int(silu_in), int(silu_out), _sp()), "small_linear")
    return out


def upsample2x(x):
    _require_cuda(x)
    assert x.is_contiguous() and x.dtype == torch.float16
    b, h, w, c = x.shape
    y = torch.empty((b, 2 * h, 2 * w, c), device=x.device, dtype=torch.float16)
    _count(1)
    check(_lib.load().ctrlora_upsample2x_f16(_dp(x), _dp(y), b, h, w, c, _sp()), "upsample2x")
    return y


def im2col_s2(x, pad_lo=1):
    """stride-2 3x3 gather; pad_lo=1: Conv2d(padding=1); pad_lo=0: zeros on the right/bottom only (the VAE's Downsample)"""
    _require_cuda(x)
    assert x.is_contiguous() and x.dtype == torch.float16
    b, h, w, c = x.shape
    y = torch.empty((b, h // 2, w // 2, 9 * c), device=x.device, dtype=torch.float16)
    _count(1)
    check(_lib.load().ctrlora_im2col_s2_pad_f16(_dp(x), _dp(y), b, h, w, c, int(pad_lo), _sp()), "im2col_s2")
    return y


def softmax_rows(logits, scale=1.0):
    """fp32 [rows, cols] -> fp16 softmax(scale * logits) over the last dim"""
    _require_cuda(logits)
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    out = torch.empty(logits.shape, device=logits.device, dtype=torch.float16)
    _count(1)
    check(_lib.load().ctrlora_softmax_rows_f32_to_f16(_dp(logits), logits.stride(0), _dp(out), out.stride(0), logits.shape[0],
                                                      logits.shape[1], float(scale), _sp()), "softmax_rows")
    return out


def gaussian_sample(moments, noise=None, scale=1.0):
    """moments fp32 [B, 2Z, H, W] -> scale * (mean + std * noise) (noise None: scale * mean), fp32 [B, Z, H, W]"""
    _require_cuda(moments, noise)
    moments = moments.float().contiguous()
    b, z2, h, w = moments.shape
    out = torch.empty((b, z2 // 2, h, w), device=moments.device, dtype=torch.float32)
    if noise is not None:
        noise = noise.float().contiguous()
        assert noise.shape == out.shape
    _count(1)
    check(_lib.load().ctrlora_gaussian_sample(_dp(moments), _dp(noise), _dp(out), b, z2 // 2, h * w, float(scale), _sp()),
          "gaussian_sample")
    return out


def cast_transpose(src, batch, rows, cols, out=None):
    """fp32 [batch, rows, cols] -> fp16 [batch, cols, rows]"""
    _require_cuda(src)
    assert src.dtype == torch.float32 and src.is_contiguous() and src.numel() == batch * rows * cols
    if out is None:
        out = torch.empty((batch, cols, rows), device=src.device, dtype=torch.float16)
    _count(1)
    check(_lib.load().ctrlora_cast_transpose_f32_to_f16(_dp(src), _dp(out), batch, rows, cols, _sp()), "cast_transpose")
    return out


def ddim_update(x, e_cond, e_uncond, cfg_scale, a_t, a_prev, sigma_t, sqrt_one_minus_at, noise=None, temperature=1.0,
                stats=None):
    """One DDIM update (cldm/ddim_hacked.py:190-231). fp32 [B,C,H,W] contiguous tensors; returns (x_prev, pred_x0)."""
    _require_cuda(x, e_cond, e_uncond, noise)
    for t in (x, e_cond, e_uncond, noise):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous() and t.shape == x.shape)
    x_prev, pred_x0 = torch.empty_like(x), torch.empty_like(x)
    b = x.shape[0]
    _count(1)
    check(_lib.load().ctrlora_ddim_update(_dp(x), _dp(e_cond), _dp(e_uncond), _dp(noise), _dp(x_prev), _dp(pred_x0),
                                          _dp(stats), b, x[0].numel(), float(cfg_scale), float(a_t), float(a_prev),
                                          float(sigma_t), float(sqrt_one_minus_at), float(temperature), _sp()),
          "ddim_update")
    return x_prev, pred_x0


def im2col_3x3(x):
    """fp16 [B,H,W,C] -> [B,H,W,9*C] (stride 1, pad 1; tap-major like the conv kernel weights)"""
    _require_cuda(x)
    assert x.is_contiguous() and x.dtype == torch.float16
    b, h, w, c = x.shape
    y = torch.empty((b, h, w, 9 * c), device=x.device, dtype=torch.float16)
    _count()
    check(_lib.load().ctrlora_im2col_3x3_f16(_dp(x), _dp(y), b, h, w, c, _sp()), "im2col_3x3")
    return y


def outer_accum(dy, x, out, alpha=1.0, beta=1.0, silu_x=False):
    """out[n, k] = beta*out + alpha * sum_b dy[b, n] f(x[b, k])  (fp32 [B,N], [B,K] -> fp32 [N,K], row strides free)"""
    _require_cuda(dy, x, out)
    assert dy.dtype == x.dtype == out.dtype == torch.float32 and dy.stride(1) == 1 and x.stride(1) == 1 and out.stride(1) == 1
    rows, n = dy.shape
    k = x.shape[1]
    assert x.shape[0] == rows and out.shape == (n, k)
    _count()
    check(_lib.load().ctrlora_outer_accum_f32(_dp(dy), dy.stride(0), _dp(x), x.stride(0), _dp(out), out.stride(0), rows, n, k,
                                              float(alpha), float(beta), int(silu_x), _sp()), "outer_accum")
    return out


def copy2d(src, dst, rows, cols, lds, ldd, accumulate=False):
    """dst[r, c] (+)= src[r, c] over fp32 [rows, cols] blocks with explicit row strides"""
    _require_cuda(src, dst)
    assert src.dtype == dst.dtype == torch.float32
    _count()
    check(_lib.load().ctrlora_copy2d_f32(_dp(src), lds, _dp(dst), ldd, rows, cols, int(accumulate), _sp()), "copy2d")
    return dst


def silu_bwd(d, x):
    """d * silu'(x), fp32"""
    _require_cuda(d, x)
    d, x = d.contiguous(), x.contiguous()
    out = torch.empty_like(d)
    _count()
    check(_lib.load().ctrlora_silu_bwd_f32(_dp(d), _dp(x), _dp(out), d.numel(), _sp()), "silu_bwd")
    return out


def cast_rows(src, rows, cols, lds):
    """fp32 [rows, cols] with row stride lds -> dense fp16 [rows, cols]"""
    _require_cuda(src)
    out = torch.empty((rows, cols), device=src.device, dtype=torch.float16)
    _count()
    check(_lib.load().ctrlora_cast_rows_f32_to_f16(_dp(src), lds, _dp(out), rows, cols, _sp()), "cast_rows")
    return out


def q_sample(x0, noise, t, tab_a, tab_s):
    """tab_a[t] * x0 + tab_s[t] * noise (fp32 [B,...] contiguous; t int64 [B]; tables fp32 on the device). Bit-exact."""
    _require_cuda(x0, noise, t, tab_a, tab_s)
    x0, noise = x0.float().contiguous(), noise.float().contiguous()
    assert t.dtype == torch.int64 and tab_a.dtype == torch.float32 and tab_s.dtype == torch.float32
    assert x0.shape == noise.shape and t.numel() == x0.shape[0]
    out = torch.empty_like(x0)
    _count()
    check(_lib.load().ctrlora_q_sample(_dp(x0), _dp(noise), _dp(t.contiguous()), _dp(tab_a.contiguous()),
                                       _dp(tab_s.contiguous()), _dp(out), x0.shape[0], x0[0].numel(), _sp()), "q_sample")
    return out


def ddim_encode_update(x, e_cond, e_uncond, cfg_scale, c1, c2):
    """x_next = c1 * x + c2 * cfg(e_cond, e_uncond)   (DDIM inversion, cldm/ddim_hacked.py:253-267)"""
    _require_cuda(x, e_cond, e_uncond)
    x, e_cond = x.float().contiguous(), e_cond.float().contiguous()
    e_uncond = None if e_uncond is None else e_uncond.float().contiguous()
    out = torch.empty_like(x)
    _count()
    check(_lib.load().ctrlora_ddim_encode_update(_dp(x), _dp(e_cond), _dp(e_uncond), _dp(out), x.numel(), float(cfg_scale),
                                                 float(c1), float(c2), _sp()), "ddim_encode_update")
    return out


def wgrad_tn(a, b, out=None, alpha=1.0, beta=0.0):
    """out[p, q] = alpha * sum_m a[m, p] * b[m, q] + beta * out  (fp16 a [M,P], b [M,Q] -> fp32 [P,Q])."""
    _require_cuda(a, b)
    assert a.dtype == torch.float16 and b.dtype == torch.float16 and a.stride(1) == 1 and b.stride(1) == 1
    m, pd = a.shape
    qd = b.shape[1]
    assert b.shape[0] == m
    if out is None:
        assert beta == 0.0
        out = torch.empty((pd, qd), device=a.device, dtype=torch.float32)
    ws, _ = _splitk_buffers(a.device)
    _count(2)
    check(_lib.load().ctrlora_wgrad_tn_f16(_dp(a), a.stride(0), _dp(b), b.stride(0), m, pd, qd, _dp(out), out.stride(0),
                                           float(alpha), float(beta), _dp(ws), SPLITK_WS_BYTES, _sp()), "ctrlora_wgrad_tn_f16")
    return out


# ------------------------------------------------------------------------------------------------ training kernels
def _gn_args(x1, gamma, beta, eps, silu, add1, add1_scale, x2, add2, add2_scale, groups, stats_ws):
    b, h, w, c1, ld1 = _as_bhwc(x1)
    c2, ld2 = 0, 0
    if x2 is not None:
        _, _, _, c2, ld2 = _as_bhwc(x2)
    a = _lib.GroupNormArgs()
    a.x1, a.add1, a.add1_scale, a.c1, a.ld1 = _dp(x1), _dp(add1), float(add1_scale), c1, ld1
    a.x2, a.add2, a.add2_scale, a.c2, a.ld2 = _dp(x2), _dp(add2), float(add2_scale), c2, ld2
    a.batch, a.hw, a.groups = b, h * w, groups
    a.gamma, a.beta, a.eps, a.silu = _dp(gamma), _dp(beta), float(eps), int(silu)
    a.stats_ws = _dp(stats_ws)
    return a, (b, h, w, c1, c2)


def groupnorm_bwd(dy, fwd_stats, x1, gamma, beta, eps, silu, *, add1=None, add1_scale=1.0, x2=None, add2=None,
                  add2_scale=1.0, groups=32, want_dx2=False, dx2_scale=1.0, dgamma=None, dbeta=None, res=None, dx1_scale=1.0):
    """Backward of ops.groupnorm (same source description).  Returns dx1 (and dx2 scaled by dx2_scale if want_dx2);
    dgamma/dbeta (fp32 [C]) are accumulated into when given."""
    _require_cuda(dy, x1)
    assert dy.dtype == torch.float16 and dy.is_contiguous()
    ws, prezeroed = _stats_take(dy.device, fwd_stats.numel())
    a, (b, h, w, c1, c2) = _gn_args(x1, gamma, beta, eps, silu, add1, add1_scale, x2, add2, add2_scale, groups, ws)
    a.stats_prezeroed = prezeroed
    dx1 = torch.empty((b, h, w, c1), device=dy.device, dtype=torch.float16)
    dx2 = torch.empty((b, h, w, c2), device=dy.device, dtype=torch.float16) if (want_dx2 and c2) else None
    _count(2)
    if res is not None:
        assert res.dtype == torch.float16 and res.stride(-1) == 1 and res.shape[-1] == c1 + c2
    check(_lib.load().ctrlora_groupnorm_bwd_f16(C.addressof(a), _dp(dy), _dp(fwd_stats), _dp(dx1), c1, float(dx1_scale), _dp(dx2), c2,
                                                float(dx2_scale), _dp(res), res.stride(-2) if res is not None else 0,
                                                _dp(dgamma), _dp(dbeta), _sp()), "groupnorm_bwd")
    return (dx1, dx2) if want_dx2 else dx1


def layernorm_bwd(x, dy, gamma, eps=1e-5, dgamma=None, dbeta=None, res=None):
    _require_cuda(x, dy)
    cols = x.shape[-1]
    x2, d2 = x.reshape(-1, cols), dy.reshape(-1, cols)
    dx = torch.empty((x2.shape[0], cols), device=x.device, dtype=torch.float16)
    _count()
    check(_lib.load().ctrlora_layernorm_bwd_f16(_dp(x2), x2.stride(0), _dp(d2), d2.stride(0), _dp(dx), cols, x2.shape[0],
                                                cols, _dp(gamma), float(eps), _dp(dgamma), _dp(dbeta), _dp(res),
                                                res.reshape(-1, cols).stride(0) if res is not None else 0, _sp()), "layernorm_bwd")
    return dx.view(x.shape)


def geglu_fwd(h):
    """h fp16 [M, 2N] = [value | gate] -> value * gelu(gate) [M, N]"""
    _require_cuda(h)
    m, n2 = h.shape
    out = torch.empty((m, n2 // 2), device=h.device, dtype=torch.float16)
    _count()
    check(_lib.load().ctrlora_geglu_fwd_f16(_dp(h), _dp(out), m, n2 // 2, _sp()), "geglu_fwd")
    return out


def geglu_bwd(h, dout):
    _require_cuda(h, dout)
    m, n2 = h.shape
    dh = torch.empty_like(h)
    _count()
    check(_lib.load().ctrlora_geglu_bwd_f16(_dp(h), _dp(dout), _dp(dh), m, n2 // 2, _sp()), "geglu_bwd")
    return dh


def colsum(x, out, scale=1.0):
    """out[c] += scale * sum_rows x[row, c]; x fp16/fp32 [rows, cols] (row stride free), out fp32 [cols]"""
    _require_cuda(x, out)
    x2 = x.reshape(-1, x.shape[-1])
    _count()
    check(_lib.load().ctrlora_colsum(_dp(x2), int(x2.dtype == torch.float32), x2.stride(0), x2.shape[0], x2.shape[1],
                                     float(scale), _dp(out), _sp()), "colsum")
    return out


def image_colsum(x, images, out):
    """out[img, c] += sum over the image's rows of x (fp16 [images*rows, cols]); out fp32 [images, >=cols] (row stride free)"""
    _require_cuda(x, out)
    x2 = x.reshape(-1, x.shape[-1])
    _count()
    check(_lib.load().ctrlora_image_colsum_f16(_dp(x2), x2.stride(0), images, x2.shape[0] // images, x2.shape[1], _dp(out),
                                               out.stride(0), _sp()), "image_colsum")
    return out


def upsample2x_bwd(dout):
    _require_cuda(dout)
    b, h2, w2, c = dout.shape
    din = torch.empty((b, h2 // 2, w2 // 2, c), device=dout.device, dtype=torch.float16)
    _count()
    check(_lib.load().ctrlora_upsample2x_bwd_f16(_dp(dout.contiguous()), _dp(din), b, h2 // 2, w2 // 2, c, _sp()), "upsample2x_bwd")
    return din


def im2col_s2_bwd(dcol, h, w):
    """dcol fp16 [B, h/2, w/2, 9*C] -> dx [B, h, w, C]"""
    _require_cuda(dcol)
    b = dcol.shape[0]
    c = dcol.shape[-1] // 9
    dx = torch.empty((b, h, w, c), device=dcol.device, dtype=torch.float16)
    _count()
    check(_lib.load().ctrlora_im2col_s2_bwd_f16(_dp(dcol.contiguous()), _dp(dx), b, h, w, c, _sp()), "im2col_s2_bwd")
    return dx


def mse_loss_grad(eps, noise, c_pad=8, grad_scale=1.0):
    """eps, noise fp32 [B,C,H,W] -> (loss fp32 [1], grad fp16 pixel-major [B,H,W,c_pad])"""
    _require_cuda(eps, noise)
    b, c, h, w = eps.shape
    loss = torch.empty(1, device=eps.device, dtype=torch.float32)
    grad = torch.empty((b, h, w, c_pad), device=eps.device, dtype=torch.float16)
    _count()
    check(_lib.load().ctrlora_mse_loss_grad(_dp(eps.contiguous()), _dp(noise.contiguous()), _dp(loss), _dp(grad), b, c, h * w,
                                            c_pad, float(grad_scale), _sp()), "mse_loss_grad")
    return loss, grad


def adamw_step(params, grads, exp_avg, exp_avg_sq, step, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
               grad_scale=1.0, skip_flag=None, bc_dev=None):
    """In-place AdamW over flat fp32 buffers (torch.optim.AdamW semantics).  skip_flag: device int32 [1]; non-zero =
    the step is skipped (non-finite gradients under loss scaling).  bc_dev: device fp32 [2] bias corrections from
    adamw_begin (then `step` is ignored)."""
    _require_cuda(params, grads)
    _count()
    check(_lib.load().ctrlora_adamw_f32(_dp(params), _dp(grads), _dp(exp_avg), _dp(exp_avg_sq), params.numel(), float(lr),
                                        float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step),
                                        float(grad_scale), _dp(skip_flag), _dp(bc_dev), _sp()), "adamw")


def adamw_begin(step_counter, skip_flag, betas, bc, skipped=None):
    """device-side step bookkeeping in front of adamw_step (see ctrlora_adamw_begin)"""
    _require_cuda(step_counter, bc)
    _count()
    check(_lib.load().ctrlora_adamw_begin(_dp(step_counter), _dp(skip_flag), float(betas[0]), float(betas[1]), _dp(bc),
                                          _dp(skipped), _sp()), "adamw_begin")


def nonfinite_flag(x, flag):
    """flag (int32 [1], device) |= any(!isfinite(x)); x fp32 flat."""
    _require_cuda(x, flag)
    assert x.dtype == torch.float32 and x.is_contiguous() and flag.dtype == torch.int32
    _count()
    check(_lib.load().ctrlora_nonfinite_flag_f32(_dp(x), x.numel(), _dp(flag), _sp()), "nonfinite_flag")
    return flag


def weighted_sum(tensors, weights, out=None):
    """sum_i weights[i] * tensors[i] over up to 8 same-shape, same-stride dense fp16 tensors (fp32 accumulate)."""
    _require_cuda(*tensors)
    t0 = tensors[0]
    n = len(tensors)
    assert 1 <= n <= 8 and len(weights) == n
    for t in tensors:
        assert t.dtype == torch.float16 and t.shape == t0.shape and t.stride() == t0.stride()
    dense = t0.is_contiguous() or t0.is_contiguous(memory_format=torch.channels_last)
    assert dense and t0.numel() % 8 == 0, "weighted_sum needs dense tensors with numel % 8 == 0"
    if out is None:
        out = torch.empty_like(t0)  # preserves the (dense) strides
    assert out.stride() == t0.stride()
    srcs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    ws = (C.c_float * n)(*[float(w) for w in weights])
    _count()
    check(_lib.load().ctrlora_weighted_sum_f16(srcs, ws, n, _dp(out), t0.numel(), _sp()), "weighted_sum")
    return out


def attention_bwd(q, k, v, o, dout, lse, batch, heads, nq, nk, head_dim, dq=None, dk=None, dv=None):
    """Backward of ops.attention.  q/o/dout [batch*nq, H*d], k/v natural [batch*nk, H*d] (fp16, row strides free),
    lse fp32 [batch, H, nq] from the forward.  Returns (dq, dk, dv) fp16."""
    _require_cuda(q, k, v, o, dout, lse)
    c = heads * head_dim
    dq = torch.empty((batch * nq, c), device=q.device, dtype=torch.float16) if dq is None else dq
    dk = torch.empty((batch * nk, c), device=q.device, dtype=torch.float16) if dk is None else dk
    dv = torch.empty((batch * nk, c), device=q.device, dtype=torch.float16) if dv is None else dv
    delta = torch.empty(batch * heads * nq, device=q.device, dtype=torch.float32)
    _count(3)
    check(_lib.load().ctrlora_attention_bwd_f16(_dp(q), q.stride(0), _dp(k), k.stride(0), _dp(v), v.stride(0), _dp(o), o.stride(0),
                                                _dp(dout), dout.stride(0), _dp(lse), _dp(delta), _dp(dq), dq.stride(0), _dp(dk),
                                                dk.stride(0), _dp(dv), dv.stride(0), batch, heads, nq, nk, head_dim, _sp()),
          "ctrlora_attention_bwd_f16")
    return dq, dk, dv


def transpose_f16(src, batch, rows, cols):
    """fp16 [batch, rows, cols] -> fp16 [batch, cols, rows]"""
    _require_cuda(src)
    assert src.dtype == torch.float16 and src.is_contiguous() and src.numel() == batch * rows * cols
    out = torch.empty((batch, cols, rows), device=src.device, dtype=torch.float16)
    _count()
    check(_lib.load().ctrlora_transpose_f16(_dp(src), _dp(out), batch, rows, cols, _sp()), "transpose_f16")
    return out


def conv_dgrad_weight(w16):
    """fp16 conv kernel weight [Cout, taps, Cin] -> data-gradient weight [Cin, taps reversed, Cout]"""
    _require_cuda(w16)
    co, taps, ci = w16.shape
    out = torch.empty((ci, taps, co), device=w16.device, dtype=torch.float16)
    _count()
    check(_lib.load().ctrlora_conv_dgrad_weight_f16(_dp(w16), _dp(out), co, taps, ci, _sp()), "conv_dgrad_weight")
    return out


def set_sm_limit(limit):
    """persistent GEMM grids use at most `limit` SMs (0 = all); baked into CUDA graphs at capture"""
    check(_lib.load().ctrlora_set_sm_limit(int(limit)), "set_sm_limit")
