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

#define CTRLORA_ABI_VERSION 1

/* status codes */
#define CTRLORA_STATUS_OK 0
#define CTRLORA_STATUS_BAD_ARGUMENT 1
#define CTRLORA_STATUS_CUDA_ERROR 2
#define CTRLORA_STATUS_TENSORMAP_ERROR 3
#define CTRLORA_STATUS_UNSUPPORTED 4

int ctrlora_abi_version(void);
/* Last CUDA error string seen by this library on the calling thread's device (host pointer, static storage). */
const char* ctrlora_last_cuda_error(void);

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
    const void* residual;   /* fp16 [M, ldr] added after scaling, or NULL */
    int ldr;
    float out_scale;        /* applied to (acc + bias + rowbias) */
    int head_dim, tok_pad;  /* for transposed stores */
    int bf16;               /* must be 0 (fp16 operands) in this ABI version */
} ctrlora_gemm_args;

int ctrlora_gemm_f16(const ctrlora_gemm_args* args, void* stream);

/* Bring-up / bisecting twin of ctrlora_gemm_f16 on the CUDA cores (same arguments, same results up to fp32
 * summation order).  Only the tests call it. */
int ctrlora_gemm_f16_simt(const ctrlora_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTRLORA_B200_H */
