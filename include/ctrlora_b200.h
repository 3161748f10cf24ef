/* ctrlora_b200 C ABI — the drop-in boundary underneath the reference's Python module contract.
 *
 * The reference (xyfJASON/ctrlora) has no FFI of its own: its hot path is torch calls.  Every entry
 * point below replaces one group of those torch call sites with a hand-written sm_100a kernel; the reference file:line
 * each one stands in for is cited on the declaration.  Conventions: plain pointers and sizes, device pointers unless
 * stated, no allocation inside (workspaces are passed in), the launch goes on `stream` (a cudaStream_t passed as
 * void*), the return value is a status code (0 = ok), no exceptions cross the boundary.
 *
 * Activations are NHWC fp16 ("pixel-major"): a [B, H, W, C] tensor is a [B*H*W, C] row-major matrix, so the
 * transformer's 'b c h w -> b (h w) c' rearranges (reference ldm/modules/attention.py:330,337) are no-ops.
 */
#ifndef CTRLORA_B200_H
#define CTRLORA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CTRLORA_ABI_VERSION 2

/* status codes */
#define CTRLORA_STATUS_OK 0
#define CTRLORA_STATUS_BAD_ARGUMENT 1
#define CTRLORA_STATUS_CUDA_ERROR 2
#define CTRLORA_STATUS_TENSORMAP_ERROR 3
#define CTRLORA_STATUS_UNSUPPORTED 4

int ctrlora_abi_version(void);
/* Last CUDA error string seen by this library on the calling thread's device (host pointer, static storage). */
const char* ctrlora_last_cuda_error(void);

/* cudaMemsetAsync(ptr, 0, bytes) on `stream`: a memset node, not a kernel (zero-initialised key padding of V^T etc.) */
int ctrlora_memset_zero(void* ptr, long long bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Implicit GEMM on the 5th-gen tensor cores (tcgen05, accumulators in TMEM, operands staged by TMA):
 *   out[m, n] = epilogue( sum_{tap, c} A[pixel(m) + tap_offset, c] * W[n, tap, c]  (+ sum_c A2[pixel(m), c] * W2[n, c]) )
 * replaces  nn.Linear / F.linear              ldm/modules/attention.py:154-161,52,72; cldm/lora.py:287-290
 *           nn.Conv2d 1x1 and 3x3 stride 1    ldm/modules/diffusionmodules/openaimodel.py:196,228-240,729; cldm/cldm.py:281-282
 *           GEGLU                             ldm/modules/attention.py:49-56      (geglu = 1)
 *           the ResBlock skip connection      openaimodel.py:233-240,272-274      (a2/w2 = 1x1 skip conv, or residual)
 *           h + emb_out[..., None, None]      openaimodel.py:263-270              (rowbias)
 *           control_i * control_scales[i]     cldm/cldm_ctrlora_finetune.py:79    (out_scale)
 * A plain [M, K] matrix is a_b = 1, a_h = 1, a_w = M, a_c = K, kh = kw = 1, pad = 0.
 */
typedef struct ctrlora_gemm_args {
    const void* a;          /* fp16 activations, pixel-major; channel stride 1, pixel stride a_ld elements */
    int a_b, a_h, a_w, a_c;
    long long a_ld;
    const void* w;          /* fp16 weights [n (2n for GEGLU), kh*kw, a_c], dense */
    int kh, kw, pad;        /* 1x1 (pad 0) or 3x3 (pad 1), stride 1 */
    const void* a2;         /* optional second activation operand with the same B,H,W (1x1), or NULL */
    int a2_c;
    long long a2_ld;
    const void* w2;         /* [n, a2_c] */
    int n;                  /* output columns */
    int block_n;            /* 0 = choose automatically */
    int geglu;
    void* out[3];           /* out[0] always; out[1..2] when seg_width > 0 */
    int seg_width;          /* 0 = single output; else column n is stored to out[n / seg_width] at column n % seg_width */
    int transposed[3];      /* segment stored as [image, seg_width, tok_pad] (i.e. [image, head, d, token]) */
    int ldc;                /* row stride (elements) of the non-transposed outputs */
    int out_f32;            /* 0: fp16 outputs, 1: fp32 outputs */
    const float* bias;      /* [n] ([2n] for GEGLU) or NULL */
    const float* rowbias;   /* [images, n] fp32 or NULL */
    int rows_per_img;       /* rows per image for rowbias / transposed stores; 0 = a_h * a_w */
    int rowbias_ld;         /* row stride of rowbias (0 = n): lets one batched time-embedding GEMV feed every ResBlock */
    const void* residual;   /* [M, ldr] added after scaling, or NULL; fp16 unless residual_f32 */
    int ldr;
    int residual_f32;
    float out_scale;        /* applied to (acc + bias + rowbias) */
    int head_dim, tok_pad;  /* for transposed stores */
    int bf16;               /* must be 0 (fp16 operands) in this ABI version */
    int split_k;            /* 0 = choose automatically (needs the workspace below), 1 = never split */
    float* splitk_ws;       /* fp32 scratch for the per-split partial tiles (no initial contents required) */
    long long splitk_ws_bytes;
    unsigned int* splitk_counters;   /* arrival counters: all zero on entry, left all zero on exit */
    int splitk_counters_len;
    void* dup_out;          /* optional: transposed segments are also stored row-major here (fp16, row stride dup_ld) */
    int dup_ld;
    int force_single_cta;   /* 1: never use the 2-CTA (cta_group::2) tile pairs (tests / bisecting) */
} ctrlora_gemm_args;

int ctrlora_gemm_f16(const ctrlora_gemm_args* args, void* stream);

/* Bring-up / bisecting twin of ctrlora_gemm_f16 on the CUDA cores (same arguments, same results up to fp32
 * summation order).  Only the tests call it. */
int ctrlora_gemm_f16_simt(const ctrlora_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * GroupNorm (+SiLU) over pixel-major fp16, fp32 statistics, optionally over the channel concatenation
 * [x1 (+ add1_scale*add1) | x2 (+ add2_scale*add2)].
 * replaces  GroupNorm32 + SiLU              ldm/modules/diffusionmodules/util.py:202-219, openaimodel.py:190-197,221-231,726-730
 *           Normalize (eps 1e-6)            ldm/modules/attention.py:88-89,327
 *           h += control.pop(); cat([h, hs.pop() + control.pop()], 1)    cldm/cldm.py:34-42
 */
typedef struct ctrlora_groupnorm_args {
    const void* x1; const void* add1; float add1_scale; int c1; long long ld1;
    const void* x2; const void* add2; float add2_scale; int c2; long long ld2;   /* x2 = NULL: single source */
    int batch, hw, groups;
    const float* gamma; const float* beta; float eps; int silu;
    void* y;          /* fp16 [batch*hw, c1+c2] */
    void* raw_out;    /* optional fp16 [batch*hw, c1+c2]: the concatenated (and summed) input itself, or NULL */
    void* stats_ws;   /* fp32 workspace [batch * groups * 2] */
    int stats_prezeroed; /* 1: the caller guarantees stats_ws is zero on entry (e.g. one memset per step over an arena of
                            workspaces): no memset node in front of the statistics kernel, which is then PDL-chained */
    float* partial_ws;   /* forward only, optional: scratch for per-block partial statistics (any contents).  With it the
                            forward is bit-reproducible: no fp32 atomics, the last block of an image sums the partials in a
                            fixed order.  NULL: round-1 behaviour (atomic accumulation into stats_ws). */
    long long partial_ws_floats;
    unsigned int* partial_counters;   /* [>= batch] arrival counters: all zero on entry, left all zero on exit */
    int partial_counters_len;
} ctrlora_groupnorm_args;
int ctrlora_groupnorm_f16(const ctrlora_groupnorm_args* args, void* stream);

/* LayerNorm over the last dim (eps 1e-5 in the reference: ldm/modules/attention.py:263-265), fp16 in/out. */
int ctrlora_layernorm_f16(const void* x, long long ldx, void* y, long long ldy, int rows, int cols,
                          const float* gamma, const float* beta, float eps, void* stream);

/* Fused attention forward: out[b, i, h*d:(h+1)*d] = softmax_j(q_i . k_j * d^-1/2) v_j   (fp32 logits / softmax).
 * replaces CrossAttention.forward   ldm/modules/attention.py:163-194 (and MemoryEfficientCrossAttention :197-243).
 * q [batch, nq, heads*d] (row stride ldq), k [batch, nk, heads*d] (ldk), vt = V transposed [batch, heads, d, nk_pad]. */
int ctrlora_attention_f16(const void* q, long long ldq, const void* k, long long ldk, const void* vt, int nk_pad,
                          void* out, long long ldo, float* lse /* optional [batch, heads, nq], log2 domain */, int batch,
                          int heads, int nq, int nk, int head_dim, void* stream);

/* Module-boundary layout/dtype conversion (the reference's tensors are NCHW fp32). */
int ctrlora_nchw_f32_to_nhwc_f16(const float* src, void* dst, int batch, int channels, int hw, int c_pad, void* stream);
int ctrlora_nhwc_to_nchw_f32(const void* src, int src_is_f32, long long ld, float* dst, int batch, int channels, int hw,
                             void* stream);

/* timestep_embedding   ldm/modules/diffusionmodules/util.py:154-174: out[b] = [cos(t_b * f) | sin(t_b * f)];
 * freqs (fp32 [half]) are computed on the host exactly as the reference does; t is int64. */
int ctrlora_timestep_embedding(const long long* t, const float* freqs, float* out, int batch, int half, void* stream);

/* y = act_out(act_in(x) W^T + b) for M = batch rows, fp32 activations, fp16 weights [n, k].
 * replaces time_embed (Linear-SiLU-Linear, openaimodel.py:526-531) and every ResBlock emb_layers (SiLU-Linear, :208-215). */
int ctrlora_small_linear(const float* x, int ldx, const void* w, const float* bias, float* y, int ldy, int rows, int n,
                         int k, int silu_in, int silu_out, void* stream);

/* F.interpolate(scale_factor=2, mode='nearest')   openaimodel.py:115 */
int ctrlora_upsample2x_f16(const void* src, void* dst, int batch, int h, int w, int channels, void* stream);
/* gather for Downsample's conv3x3 stride 2 pad 1 (openaimodel.py:148-159): dst [batch, h/2, w/2, 9, channels] */
int ctrlora_im2col_s2_f16(const void* src, void* dst, int batch, int h, int w, int channels, void* stream);
/* same gather with the zero padding only on the right/bottom when pad_lo = 0: the first-stage VAE's
 * F.pad(x, (0,1,0,1)) + Conv2d(stride 2, padding 0), ldm/modules/diffusionmodules/model.py:80-84 */
int ctrlora_im2col_s2_pad_f16(const void* src, void* dst, int batch, int h, int w, int channels, int pad_lo, void* stream);
/* softmax(scale * src) over rows, fp32 logits -> fp16 probabilities: the VAE's d = 512 single-head AttnBlock
 * (model.py:179-203), whose logits come from ctrlora_gemm_f16 with out_f32 = 1 */
int ctrlora_softmax_rows_f32_to_f16(const float* src, long long lds, void* dst, long long ldd, long long rows, int cols,
                                    float scale, void* stream);
/* DiagonalGaussianDistribution.sample() / .mode() times scale_factor (ldm/modules/distributions/distributions.py:24-37,
 * ldm/models/diffusion/ddpm.py get_first_stage_encoding): moments fp32 [batch, 2*z, hw]; noise NULL = mode. */
int ctrlora_gaussian_sample(const float* moments, const float* noise, float* out, int batch, int z_channels, int hw, float scale,
                            void* stream);
/* weight preparation: fp32 [batch, rows, cols] -> fp16 [batch, cols, rows] */
int ctrlora_cast_transpose_f32_to_f16(const float* src, void* dst, long long batch, int rows, int cols, void* stream);

/* fp16 [batch, rows, cols] -> fp16 [batch, cols, rows] (transposed weight copies for the data-gradient GEMMs) */
int ctrlora_transpose_f16(const void* src, void* dst, long long batch, int rows, int cols, void* stream);

/* conv kernel weight fp16 [cout, taps, cin] -> data-gradient weight [cin, taps reversed, cout]: dx = conv(dy, W_d) with the
 * same padding -- the adjoint of torch.nn.Conv2d (ldm/modules/diffusionmodules/util.py:224 conv_nd) that autograd applies;
 * rebuilt every step in pretraining, where the conv weights train (cldm/cldm_ctrlora_pretrain.py:88-96). */
int ctrlora_conv_dgrad_weight_f16(const void* src, void* dst, int cout, int taps, int cin, void* stream);

/* Persistent GEMM grids use at most `limit` SMs from now on (0 = all of them): the gradient all-reduce that overlaps the
 * ControlNet backward (the reference's DDP does the same overlap by buckets, pytorch_lightning strategy "ddp",
 * train_ctrlora_pretrain.py) owns a few SMs, and a 148-CTA persistent grid would wait for them.  Read at launch time. */
int ctrlora_set_sm_limit(int limit);

/* DDIM update in one pass   cldm/ddim_hacked.py:190-192 (CFG, e_uncond may be NULL), :208-231 (pred_x0, x_prev).
 * fp32, round-to-nearest ops in the reference's order. stats (optional, [batch]) receives sum(x_prev^2) per image. */
int ctrlora_ddim_update(const float* x, const float* e_cond, const float* e_uncond, const float* noise, float* x_prev,
                        float* pred_x0, float* stats, int batch, int per_image, float cfg_scale, float a_t, float a_prev,
                        float sigma_t, float sqrt_one_minus_at, float temperature, void* stream);

/* q_sample (ldm/models/diffusion/ddpm.py:356-359) and DDIMSampler.stochastic_encode (cldm/ddim_hacked.py:281-296):
 * out[b] = tab_a[t[b]] * x0[b] + tab_s[t[b]] * noise[b]; t int64 [batch] (device), tables fp32 (device).  Bit-exact. */
int ctrlora_q_sample(const float* x0, const float* noise, const long long* t, const float* tab_a, const float* tab_s,
                     float* out, int batch, int per_image, void* stream);
/* DDIM inversion step (cldm/ddim_hacked.py:253-267): e = e_uncond + cfg*(e_cond - e_uncond) (e_uncond may be NULL), then
 * x_next = c1 * x + c2 * e with c1 = sqrt(a_next/a), c2 = sqrt(a_next) * (sqrt(1/a_next - 1) - sqrt(1/a - 1)) evaluated
 * by the caller in fp32 like the reference's 0-dim tensor arithmetic. */
int ctrlora_ddim_encode_update(const float* x, const float* e_cond, const float* e_uncond, float* x_next, int total,
                               float cfg_scale, float c1, float c2, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Training (backward of the trainable set; reference: autograd over cldm/lora.py:70-80,285-291 and cldm/cldm.py:281-282,
 * parameters selected by cldm/cldm_ctrlora_finetune.py:88-100).
 *
 * Weight-gradient GEMM over the token dimension: out[p, q] = alpha * sum_m a[m, p] * b[m, q] + beta * out[p, q],
 * fp16 a [m, p_dim] (row stride lda), b [m, q_dim] (ldb), fp32 out (row stride ldo).  ws: fp32 scratch (any contents).
 */
int ctrlora_wgrad_tn_f16(const void* a, long long lda, const void* b, long long ldb, int m, int p_dim, int q_dim, float* out,
                         long long ldo, float alpha, float beta, float* ws, long long ws_bytes, void* stream);

/* Attention backward (autograd of ldm/modules/attention.py:163-194): v is the NATURAL [batch*nk, heads*d] layout; lse is
 * what ctrlora_attention_f16 wrote; delta_ws: fp32 scratch [batch*heads*nq].  dq/dk/dv: fp16, same layouts as q/k/v. */
int ctrlora_attention_bwd_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                              const void* o, long long ldo, const void* dout, long long lddo, const float* lse,
                              float* delta_ws, void* dq, long long lddq, void* dk, long long lddk, void* dv, long long lddv,
                              int batch, int heads, int nq, int nk, int head_dim, void* stream);

/* GroupNorm(+SiLU) backward: same source description as the forward (`args`, whose stats_ws is a scratch buffer for the
 * backward statistics); fwd_stats = the {sum, sumsq} buffer the forward left in ITS stats_ws.  dx1 / dx2: gradients of
 * the two concat halves (fp16, row strides ldd1 / ldd2, scaled by dx*_scale; dx2 may be NULL).  dgamma/dbeta (fp32 [C],
 * accumulated into) may be NULL. */
int ctrlora_groupnorm_bwd_f16(const ctrlora_groupnorm_args* args, const void* dy, const void* fwd_stats, void* dx1,
                              long long ldd1, float dx1_scale, void* dx2, long long ldd2, float dx2_scale,
                              const void* res /* optional fp16 [rows, ldres]: added to the concat gradient */, long long ldres,
                              float* dgamma, float* dbeta, void* stream);
/* LayerNorm backward (statistics recomputed from x); dgamma/dbeta accumulated into (may be NULL). */
int ctrlora_layernorm_bwd_f16(const void* x, long long ldx, const void* dy, long long ldy, void* dx, long long lddx, int rows,
                              int cols, const float* gamma, float eps, float* dgamma, float* dbeta,
                              const void* res /* optional fp16 residual-branch gradient added to dx */, long long ldres,
                              void* stream);
/* GEGLU on the stored projection h = [value | gate] ([rows, 2n]): out = value * gelu(gate), and its backward. */
int ctrlora_geglu_fwd_f16(const void* h, void* out, long long rows, int n, void* stream);
int ctrlora_geglu_bwd_f16(const void* h, const void* dout, void* dh, long long rows, int n, void* stream);
/* out[c] += scale * sum_rows x[row, c]  (bias gradients); out[img, c] += per-image column sums (time-embedding gradients) */
int ctrlora_colsum(const void* x, int x_is_f32, long long ld, long long rows, int cols, float scale, float* out, void* stream);
int ctrlora_image_colsum_f16(const void* x, long long ld, int images, int rows_per_img, int cols, float* out, long long ldo,
                             void* stream);
/* Pretraining (every ControlNet parameter trainable, cldm/cldm_ctrlora_pretrain.py:174-182): dense weight gradients.
 * dW[Cout, tap, Cin] of a 3x3 stride-1 conv = ctrlora_wgrad_tn_f16(dY [M, Cout], col [M, 9*Cin]) with
 * col[b, h, w, tap, c] = x[b, h+kh-1, w+kw-1, c]: */
int ctrlora_im2col_3x3_f16(const void* src, void* dst, int batch, int h, int w, int channels, void* stream);
/* out[n, k] = beta*out + alpha * sum_b dy[b, n] * f(x[b, k]) (fp32; b = batch rows; f = SiLU when silu_x): time_embed /
 * emb_layers weight gradients */
int ctrlora_outer_accum_f32(const float* dy, int lddy, const float* x, int ldx, float* out, long long ldo, int rows, int n, int k,
                            float alpha, float beta, int silu_x, void* stream);
/* dst[r, c] (+)= src[r, c], fp32, row strides lds / ldd */
int ctrlora_copy2d_f32(const float* src, long long lds, float* dst, long long ldd, long long rows, int cols, int accumulate,
                       void* stream);
/* out = d * silu'(x) (fp32): backward of the SiLU in the time-embedding MLP (openaimodel.py:526-531, :208-215) */
int ctrlora_silu_bwd_f32(const float* d, const float* x, float* out, long long n, void* stream);
/* fp32 [rows, cols] (row stride lds) -> dense fp16 [rows, cols] */
int ctrlora_cast_rows_f32_to_f16(const float* src, long long lds, void* dst, long long rows, int cols, void* stream);
/* adjoints of ctrlora_upsample2x_f16 and ctrlora_im2col_s2_f16 */
int ctrlora_upsample2x_bwd_f16(const void* dout, void* din, int batch, int h, int w, int channels, void* stream);
int ctrlora_im2col_s2_bwd_f16(const void* dcol, void* dx, int batch, int h, int w, int channels, void* stream);
/* loss = mean((eps - noise)^2)  (ldm/models/diffusion/ddpm.py:902-918, logvar = 0) and its gradient, written as
 * pixel-major fp16 [batch, hw, c_pad] (times grad_scale); eps / noise are fp32 NCHW. */
int ctrlora_mse_loss_grad(const float* eps, const float* noise, float* loss, void* grad, int batch, int channels, int hw,
                          int c_pad, float grad_scale, void* stream);
/* torch.optim.AdamW step over one flat fp32 buffer (cldm/cldm_ctrlora_finetune.py:105: lr 1e-5, betas .9/.999, eps 1e-8,
 * weight decay 0.01); grads are multiplied by grad_scale first (1/(world_size * loss_scale) after an all-reduce SUM).
 * skip_flag (device int, may be NULL): when non-zero the step is skipped (loss-scale overflow, see below). */
int ctrlora_adamw_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                      float beta2, float eps, float weight_decay, int step, float grad_scale, const int* skip_flag,
                      const float* bc_dev /* optional device {1-beta1^step, 1-beta2^step} from ctrlora_adamw_begin; then `step`
                                             is ignored */,
                      void* stream);
/* In front of an AdamW step: ++*step_counter unless *skip_flag (then ++*skipped), bc[0..1] = 1 - beta^step.  Keeps torch's
 * per-parameter `step` semantics (skipped steps do not count) without a host read of the overflow flag every step. */
int ctrlora_adamw_begin(int* step_counter, const int* skip_flag, float beta1, float beta2, float* bc, int* skipped, void* stream);
/* *flag |= 1 if any element of x is NaN/Inf: the overflow check of the loss-scaled fp16 backward (the reference trains in
 * fp32 and has no such step; torch.cuda.amp.GradScaler semantics: skip the update, lower the scale). */
int ctrlora_nonfinite_flag_f32(const float* x, long long n, int* flag, void* stream);
/* out = sum_i weights[i] * srcs[i] over `count` (<= 8) fp16 tensors of n elements (n % 8 == 0), fp32 accumulation:
 * the weighted control sum of multi-LoRA inference, cldm/cldm_ctrlora_inference.py:172-176.  srcs / weights: HOST arrays. */
int ctrlora_weighted_sum_f16(const void* const* srcs, const float* weights, int count, void* out, long long n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTRLORA_B200_H */
