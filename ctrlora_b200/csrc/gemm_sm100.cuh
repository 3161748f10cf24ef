// Kernel-side parameter block of the tcgen05 implicit-GEMM kernel (see gemm_sm100.cu).
#pragma once
#include "common.cuh"

namespace ctrl {

constexpr int GEMM_BM = 128;            // UMMA M (rows of the accumulator = TMEM lanes)
constexpr int GEMM_BK = 64;             // 64 fp16 = 128 B = one SWIZZLE_128B row
constexpr int GEMM_MAX_STAGES = 10;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KiB
constexpr int GEMM_EPI_WARPS = 8;       // two warps per TMEM lane group, interleaved over 32-column chunks
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;  // warp 0: TMA producer, warp 1: MMA issuer, then epilogue
constexpr int GEMM_SMEM_DATA = 216 * 1024;            // ring buffer budget
constexpr int GEMM_SMEM_BYTES = GEMM_SMEM_DATA + 1024 /*align slack*/ + 512 /*barriers*/ + 4096 /*bias staging*/;

// Tensor maps of the TMA-staged epilogue: per output segment a 64-column SWIZZLE_128B map and a 32-column SWIZZLE_64B
// map (32-row sub-box stores), the same pair for the natural-layout duplicate of a transposed segment, and the
// whole-tile pair of the fp16 residual.
struct GemmEpiMaps {
    CUtensorMap out[3][2];
    CUtensorMap dup[2];
    CUtensorMap res[2];
};

struct GemmKParams {
    // tile geometry over the (B, H, W) pixel grid; a plain [M, K] matrix is B=1, H=1, W=M
    int W, H, Bn;
    int bw, bh, nb;                 // TMA box of one 128-row tile: nb * bh * bw == 128
    int tiles_w, tiles_h, tiles_b;  // ceil-div of the dims above by the box
    int N;                          // output columns
    int BN;                         // UMMA N of one tile (GEGLU: value half + gate half)
    int n_tiles;
    int taps, kw, pad;              // filter taps (1 or 9), filter width, zero padding
    int kchunks;                    // ceil(Cin / 64) per tap
    int kchunks2;                   // extra 1x1 segment from the second operand pair (0 = none)
    int geglu;                      // 1: weights are [2N, K]; out = value * gelu(gate)
    int stages, stage_bytes;        // TMA ring: stage = A tile (16 KiB) + B tile (BN * 128 B, 1 KiB aligned)
    int splits, kiters_per_split;   // split-K: partial sums meet in `ws` (fp32, self-cleaning), last CTA runs the epilogue
    float* ws;
    unsigned int* counters;
    uint32_t idesc;
    // epilogue
    void* out[3];
    int seg_width;                  // 0: single output; else column n goes to out[n / seg_width]
    int transposed[3];              // store segment as [img, head, d, tok_pad] (V^T for attention)
    int ldc;
    int out_f32;
    const float* bias;              // [N] (GEGLU: [2N])
    const float* rowbias;           // [images, rowbias_ld]  per-image additive term (time embedding)
    int rows_per_img;
    int rowbias_ld;
    int residual_f32;
    const __half* residual;         // [M, ldr]
    int ldr;
    float out_scale;
    int head_dim, tok_pad;
    __half* dup_out;               // transposed segments are ALSO stored row-major here (training keeps natural V)
    int dup_ld;
    int bf16;
    // TMA-staged epilogue (short-K shapes): tile -> swizzled shared staging -> cp.async.bulk.tensor stores; the fp16
    // residual tile is TMA-loaded into the same staging buffer one tile ahead and updated in place.
    int epi_tma, epi_nbuf, epi_res;
    int epi_off, epi_buf_bytes;     // staging buffers live at the top of the ring area
    int epi_nfull, epi_tail;        // 64-column SWIZZLE_128B blocks (16 KiB each) + optional 32-column SWIZZLE_64B tail (8 KiB)
};

}  // namespace ctrl
