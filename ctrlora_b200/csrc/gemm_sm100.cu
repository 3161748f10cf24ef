// tcgen05 implicit-GEMM for sm_100a: one persistent, warp-specialised kernel that serves
//   * every nn.Linear of the hot path            (reference: ldm/modules/attention.py:154-161, cldm/lora.py:285-291)
//   * every 1x1 / 3x3 stride-1 Conv2d in NHWC    (reference: ldm/modules/diffusionmodules/openaimodel.py:162-274)
// D[M, N] = sum_taps A_shifted[M, Cin] * W[N, tap, Cin]^T  (+ optional second 1x1 operand pair: the ResBlock skip conv)
// A tiles are TMA boxes over the (C, W, H, B) activation tensor: a filter tap is a coordinate shift and the conv zero
// padding is the TMA out-of-bounds fill, so no im2col buffer exists in HBM.  Accumulators live in TMEM (two 256-column
// stages, so the epilogue of tile i overlaps the MMA main loop of tile i+1).
#include "gemm_sm100.cuh"
#include "ctrlora_b200.h"
#include <stdio.h>
#include <string.h>

namespace ctrl {

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// One 32-column chunk of the epilogue for one accumulator row: bias, GEGLU, time-embedding row term, scale, residual,
// then the store (row-major fp16 / fp32, or the transposed V^T layout).  v = value columns, g = gate columns (GEGLU).
template <bool GEGLU>
__device__ __forceinline__ void epilogue_chunk(const GemmKParams& p, float* v, const float* g, int c, int bn_out, int n0,
                                               bool row_ok, long long m, int img, int tok, const float* sb,
                                               const uint4* rpre = nullptr) {
    const int nbase = n0 + c;
    const bool full_chunk = (c + 32 <= bn_out) && (nbase + 32 <= p.N);
    // sb: this tile's bias staged in shared memory (zeros where there is no bias / beyond N): broadcast 16-byte reads
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 b4 = lds128f(smem_u32(sb + c + 4 * q));
        v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
    }
    if (GEGLU) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 b4 = lds128f(smem_u32(sb + 256 + c + 4 * q));
            v[4 * q] *= gelu_erf_f(g[4 * q] + b4.x);
            v[4 * q + 1] *= gelu_erf_f(g[4 * q + 1] + b4.y);
            v[4 * q + 2] *= gelu_erf_f(g[4 * q + 2] + b4.z);
            v[4 * q + 3] *= gelu_erf_f(g[4 * q + 3] + b4.w);
        }
    }
    if (!row_ok) return;
    if (p.rowbias) {
        const float* rb = p.rowbias + static_cast<long long>(img) * p.rowbias_ld + nbase;
        if (full_chunk && (p.rowbias_ld & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(rb) + q);
                v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (nbase + j < p.N) v[j] += __ldg(rb + j);
        }
    }
    if (p.out_scale != 1.0f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
    }
    if (p.residual && p.residual_f32) {
        const float* rp = reinterpret_cast<const float*>(p.residual) + m * p.ldr + nbase;
        if (full_chunk && (p.ldr & 3) == 0) {  // LoRA folds: W (fp32 master) + s * up . down
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 r4 = __ldg(reinterpret_cast<const float4*>(rp) + q);
                v[4 * q] += r4.x; v[4 * q + 1] += r4.y; v[4 * q + 2] += r4.z; v[4 * q + 3] += r4.w;
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (c + j < bn_out && nbase + j < p.N) v[j] += rp[j];
        }
    } else if (p.residual) {
        const __half* rp = p.residual + m * p.ldr + nbase;
        if (full_chunk && (p.ldr & 7) == 0) {
            uint4 u[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) u[q] = rpre ? rpre[q] : __ldg(reinterpret_cast<const uint4*>(rp) + q);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const __half2* h = reinterpret_cast<const __half2*>(&u[q]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(h[e]);
                    v[q * 8 + e * 2] += f.x;
                    v[q * 8 + e * 2 + 1] += f.y;
                }
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (c + j < bn_out && nbase + j < p.N) v[j] += __half2float(rp[j]);
        }
    }
    int seg = 0, nloc = nbase;
    if (p.seg_width > 0) { seg = nbase / p.seg_width; nloc = nbase - seg * p.seg_width; }
    if (p.transposed[seg]) {
        __half* o = reinterpret_cast<__half*>(p.out[seg]) + (static_cast<long long>(img) * p.seg_width + nloc) * p.tok_pad + tok;
        for (int j = 0; j < 32; ++j)
            if (c + j < bn_out && nbase + j < p.N) o[static_cast<long long>(j) * p.tok_pad] = __float2half_rn(v[j]);
        if (p.dup_out) {
            __half* o2 = p.dup_out + m * p.dup_ld + nloc;
            for (int j = 0; j < 32; ++j)
                if (c + j < bn_out && nbase + j < p.N) o2[j] = __float2half_rn(v[j]);
        }
    } else if (p.out_f32) {
        float* o = reinterpret_cast<float*>(p.out[seg]) + m * p.ldc + nloc;
        if (full_chunk && (p.ldc & 3) == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                reinterpret_cast<float4*>(o)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {
            for (int j = 0; j < 32; ++j)
                if (c + j < bn_out && nbase + j < p.N) o[j] = v[j];
        }
    } else {
        __half* o = reinterpret_cast<__half*>(p.out[seg]) + m * p.ldc + nloc;
        if (full_chunk && (p.ldc & 7) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 u;
                u.x = pack_h2(v[q * 8 + 0], v[q * 8 + 1]);
                u.y = pack_h2(v[q * 8 + 2], v[q * 8 + 3]);
                u.z = pack_h2(v[q * 8 + 4], v[q * 8 + 5]);
                u.w = pack_h2(v[q * 8 + 6], v[q * 8 + 7]);
                reinterpret_cast<uint4*>(o)[q] = u;
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (c + j < bn_out && nbase + j < p.N) o[j] = __float2half_rn(v[j]);
        }
    }
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <bool GEGLU, bool PAIR>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const __grid_constant__ GemmEpiMaps em, const __grid_constant__ GemmKParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_SMEM_DATA);
    uint64_t* full = bars;
    uint64_t* empty = bars + GEMM_MAX_STAGES;
    uint64_t* tfull = bars + 2 * GEMM_MAX_STAGES;
    uint64_t* tempty = bars + 2 * GEMM_MAX_STAGES + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_MAX_STAGES + 4);
    volatile int* last_flag = reinterpret_cast<volatile int*>(bars + 2 * GEMM_MAX_STAGES + 5);
    uint64_t* rfull = bars + 2 * GEMM_MAX_STAGES + 6;    // residual tile landed in staging buffer i (TMA epilogue)
    uint64_t* rempty = bars + 2 * GEMM_MAX_STAGES + 8;   // staging buffer i drained by all epilogue warps
    float* sbias_all = reinterpret_cast<float*>(smem + GEMM_SMEM_DATA + 512);  // [2 tile parities][value 256 | gate 256]

    pdl_launch_dependents();
    const int warp = uniform_warp_idx();
    const int lane = threadIdx.x & 31;
    const int nstages = p.stages;
    // PAIR: two CTAs (a cluster) own one 256-row tile: tcgen05.mma.cta_group::2 issued by CTA 0 reads A (128 rows) and
    // half of B (BN/2 rows) from EACH CTA's shared memory, so every SM ingests half the B bytes per flop.
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.epi_tma) {
            tma_prefetch_desc(&em.out[0][0]);
            tma_prefetch_desc(&em.out[0][1]);
            if (p.epi_res) { tma_prefetch_desc(&em.res[0]); tma_prefetch_desc(&em.res[1]); }
        }
        if (p.kchunks2 > 0) {
            tma_prefetch_desc(&tmA2);
            tma_prefetch_desc(&tmB2);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < nstages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], PAIR ? 2 * GEMM_EPI_WARPS : GEMM_EPI_WARPS);
            mbar_init(&rfull[i], 1);
            mbar_init(&rempty[i], GEMM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) tmem_alloc_2cta(tmem_ptr, 512);
        else tmem_alloc(tmem_ptr, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / TMA completion targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();  // everything above overlapped the previous kernel's tail; operands are read only from here on

    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    const int m_units = PAIR ? (m_tiles + 1) / 2 : m_tiles;  // scheduling units along M (pairs of 128-row tiles)
    const int total_tiles = m_units * p.n_tiles * p.splits;
    const int tile_first = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int tile_stride = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int main_iters = p.taps * p.kchunks;
    const int k_iters = main_iters + p.kchunks2;
    const int bn_out = GEGLU ? (p.BN >> 1) : p.BN;

    // tile -> (m tile, split, n tile); the k range of a split is [ks * kiters_per_split, ...)
    auto decode = [&](int tile, int& mt, int& ks, int& nt) {
        mt = tile % m_units;
        if (PAIR) mt = 2 * mt + static_cast<int>(rank);  // an odd tile count leaves a phantom tile: all-OOB loads, masked rows
        const int rest = tile / m_units;
        ks = rest % p.splits;
        nt = rest / p.splits;
    };

    if (warp == 0) {
        {
            // ------------------------------------------------ TMA producer (whole warp loops; one elected lane issues)
            // One thread; its loop body is the serial critical path of the whole pipeline (measured: the first version
            // spent ~650 cycles per k-step in integer div/mod and address math, which capped every conv at ~40 % of the
            // tensor pipe regardless of tile width). Everything is strength-reduced to counters and raw shared addresses.
            const uint32_t tx_bytes = PAIR ? 2 * (GEMM_A_BYTES + (p.BN >> 1) * 128) : GEMM_A_BYTES + p.BN * 128;
            const int kchunks = p.kchunks, kw = p.kw, pad = p.pad, kps = p.kiters_per_split;
            const uint32_t stage_bytes = static_cast<uint32_t>(p.stage_bytes);
            const uint32_t smem0 = smem_u32(smem), full0 = smem_u32(full), empty0 = smem_u32(empty);
            const uint32_t gate_off = GEMM_A_BYTES + static_cast<uint32_t>(bn_out) * 128;
            const uint32_t rfull0 = smem_u32(rfull), rempty0 = smem_u32(rempty);
            int stage = 0, rbuf = 0;
            uint32_t phase = 0, a_dst = smem0, fb = full0, eb = empty0, rphase = 0;
            for (int tile = tile_first; tile < total_tiles; tile += tile_stride) {
                int mt, ks, nt;
                decode(tile, mt, ks, nt);
                const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
                const int w0 = tw * p.bw - pad, h0 = th * p.bh - pad, b0 = tb * p.nb, n0 = nt * bn_out;
                // this CTA's B rows. PAIR: GEGLU -> CTA 0 value rows, CTA 1 gate rows; else the two N halves
                const int brow = !PAIR ? n0 : GEGLU ? (leader ? n0 : p.N + n0) : n0 + static_cast<int>(rank) * (p.BN >> 1);
                const int it0 = ks * kps, it1 = min(k_iters, it0 + kps);
                int tap = 0, kc = 0, kx = 0, ky = 0;
                if (it0 > 0 && it0 < main_iters) {
                    tap = it0 / kchunks; kc = it0 - tap * kchunks;
                    ky = tap / kw; kx = tap - ky * kw;
                }
                int c0 = kc * GEMM_BK;
                const int it_main = min(it1, main_iters);
                int it = it0;
                for (; it < it_main; ++it) {
                    mbar_wait_a(eb, phase ^ 1);
                    if (elect_one()) {
                        if (!PAIR || leader) mbar_expect_tx_a(fb, tx_bytes);  // PAIR: both CTAs' bytes land on CTA 0's barrier
                        tma_load_4d_a<PAIR>(a_dst, &tmA, fb, c0, w0 + kx, h0 + ky, b0);
                        tma_load_3d_a<PAIR>(a_dst + GEMM_A_BYTES, &tmB, fb, c0, tap, brow);
                        if (GEGLU && !PAIR) tma_load_3d_a<false>(a_dst + gate_off, &tmB, fb, c0, tap, p.N + n0);
                    }
                    __syncwarp();
                    c0 += GEMM_BK;
                    if (++kc == kchunks) {
                        kc = 0; c0 = 0; ++tap;
                        if (++kx == kw) { kx = 0; ++ky; }
                    }
                    a_dst += stage_bytes; fb += 8; eb += 8;
                    if (++stage == nstages) { stage = 0; phase ^= 1; a_dst = smem0; fb = full0; eb = empty0; }
                }
                c0 = (max(it0, main_iters) - main_iters) * GEMM_BK;
                for (; it < it1; ++it) {  // second operand pair (fused 1x1 skip convolution)
                    mbar_wait_a(eb, phase ^ 1);
                    if (elect_one()) {
                        if (!PAIR || leader) mbar_expect_tx_a(fb, tx_bytes);
                        tma_load_4d_a<PAIR>(a_dst, &tmA2, fb, c0, w0 + pad, h0 + pad, b0);
                        tma_load_3d_a<PAIR>(a_dst + GEMM_A_BYTES, &tmB2, fb, c0, 0, brow);
                    }
                    __syncwarp();
                    c0 += GEMM_BK;
                    a_dst += stage_bytes; fb += 8; eb += 8;
                    if (++stage == nstages) { stage = 0; phase ^= 1; a_dst = smem0; fb = full0; eb = empty0; }
                }
                if (p.epi_res) {
                    // residual tile of THIS tile -> staging buffer; it is needed only when the tile's MMAs are done
                    mbar_wait_a(rempty0 + 8 * rbuf, rphase ^ 1);
                    if (elect_one()) {
                        const uint32_t rb = rfull0 + 8 * rbuf;
                        const uint32_t dst = smem0 + p.epi_off + rbuf * p.epi_buf_bytes;
                        mbar_expect_tx_a(rb, GEMM_BM * bn_out * 2);
                        for (int b = 0; b < p.epi_nfull; ++b)
                            tma_load_4d_a<false>(dst + b * 16384, &em.res[0], rb, n0 + 64 * b, w0 + pad, h0 + pad, b0);
                        if (p.epi_tail)
                            tma_load_4d_a<false>(dst + p.epi_nfull * 16384, &em.res[1], rb, n0 + 64 * p.epi_nfull, w0 + pad, h0 + pad, b0);
                    }
                    __syncwarp();
                    if (++rbuf == p.epi_nbuf) { rbuf = 0; rphase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {
            // ------------------------------------------------ MMA issuer (whole warp loops, one elected lane issues; CTA 0
            // issues for the pair)
            const int kps = p.kiters_per_split;
            const uint32_t idesc = p.idesc;
            const uint32_t full0 = smem_u32(full), empty0 = smem_u32(empty), tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
            // descriptors differ between stages only in the 14-bit start-address field (16-byte units)
            const uint64_t a_desc0 = umma_desc_kmajor_sw128(smem_u32(smem));
            const uint64_t b_desc0 = umma_desc_kmajor_sw128(smem_u32(smem) + GEMM_A_BYTES);
            const uint32_t stage_step = static_cast<uint32_t>(p.stage_bytes) >> 4;
            int stage = 0;
            uint32_t phase = 0, fb = full0, eb = empty0, desc_off = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = tile_first; tile < total_tiles; tile += tile_stride) {
                const int ks = (tile / m_units) % p.splits;
                const int it0 = ks * kps, it1 = min(k_iters, it0 + kps);
                mbar_wait_a(tempty0 + 8 * acc, acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                uint32_t accum = 0;
                for (int it = it0; it < it1; ++it) {
                    mbar_wait_a(fb, phase);
                    tc_fence_after();
                    const uint64_t a_desc = a_desc0 + desc_off, b_desc = b_desc0 + desc_off;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            if (PAIR) umma_f16_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (accum | k) ? 1u : 0u);
                            else umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (accum | k) ? 1u : 0u);
                        }
                        umma_commit_a<PAIR>(eb);  // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
                        // accumulator ready: committed by the SAME lane that issued the MMAs (commit tracks the issuing thread)
                        if (it + 1 == it1) umma_commit_a<PAIR>(tfull0 + 8 * acc);
                    }
                    __syncwarp();
                    accum = 1;
                    fb += 8; eb += 8; desc_off += stage_step;
                    if (++stage == nstages) { stage = 0; phase ^= 1; fb = full0; eb = empty0; desc_off = 0; }
                }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ---------------------------------------------------- epilogue warps (TMEM -> registers -> global)
        const int lane_grp = warp & 3;             // TMEM lanes [32*lane_grp, +32) are the ones this warp may read
        const int half = (warp - 2) >> 2;          // the two warps of a lane group take alternate 32-column chunks
        const int r = lane_grp * 32 + lane;
        const int iw = r % p.bw, ih = (r / p.bw) % p.bh, ib = r / (p.bw * p.bh);
        int acc = 0;
        uint32_t acc_phase = 0;
        int tile_iter = 0;
        for (int tile = tile_first; tile < total_tiles; tile += tile_stride, ++tile_iter) {
            int mt, ks, nt;
            decode(tile, mt, ks, nt);
            const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, tb = mt / (p.tiles_w * p.tiles_h);
            const int gw = tw * p.bw + iw, gh = th * p.bh + ih, gb = tb * p.nb + ib;
            const bool row_ok = gw < p.W && gh < p.H && gb < p.Bn;
            const long long m = (static_cast<long long>(gb) * p.H + gh) * p.W + gw;
            const int n0 = nt * bn_out;
            const int img = row_ok ? static_cast<int>(m / p.rows_per_img) : 0;
            const int tok = row_ok ? static_cast<int>(m % p.rows_per_img) : 0;

            // ---- stage this tile's bias in shared memory (double-buffered by tile parity)
            float* sb = sbias_all + (tile_iter & 1) * 512;
            {
                const int e = threadIdx.x - 64;  // 0..255 over the epilogue warps
                float bv = 0.f, bg = 0.f;
                if (p.bias && e < bn_out && n0 + e < p.N) {
                    bv = __ldg(p.bias + n0 + e);
                    if (GEGLU) bg = __ldg(p.bias + p.N + n0 + e);
                }
                sts32f(smem_u32(sb + e), bv);
                if (GEGLU) sts32f(smem_u32(sb + 256 + e), bg);
            }
            // ---- TMA epilogue: make sure this tile's staging buffer is free (our own bulk stores of `nbuf` tiles ago
            // have finished reading it); with a residual the buffer is handed back to the producer warp instead
            const int sbuf = p.epi_nbuf == 2 ? (tile_iter & 1) : 0;
            const uint32_t sphase = p.epi_nbuf == 2 ? ((tile_iter >> 1) & 1) : (tile_iter & 1);
            if (p.epi_tma) {
                if (lane == 0) {
                    if (p.epi_res) {
                        if (tile_iter > 0) {
                            bulk_wait_group_read<0>();
                            mbar_arrive(&rempty[p.epi_nbuf == 2 ? ((tile_iter - 1) & 1) : 0]);
                        }
                    } else if (p.epi_nbuf == 2) {
                        bulk_wait_group_read<1>();
                    } else {
                        bulk_wait_group_read<0>();
                    }
                }
                __syncwarp();
            }
            asm volatile("bar.sync 2, %0;" ::"n"(32 * GEMM_EPI_WARPS) : "memory");  // bias staged
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(lane_grp * 32) << 16);

            if (p.epi_tma) {
                // ---- TMEM -> registers -> swizzled staging (in place over the residual tile) -> TMA store per warp
                if (p.epi_res) mbar_wait(&rfull[sbuf], sphase);
                uint8_t* stg = smem + p.epi_off + sbuf * p.epi_buf_bytes;
                const uint32_t sb_a = smem_u32(sb);
                const int nblk = p.epi_nfull + p.epi_tail;
                // output segment of this tile (q | k | v projections share one GEMM); a transposed segment (V^T for the
                // attention kernels) is stored straight from registers -- thread = token makes THAT store the coalesced
                // one -- and only its natural-layout duplicate (training) goes through the staging buffer
                const int seg = p.seg_width > 0 ? n0 / p.seg_width : 0;
                const int nloc0 = n0 - seg * p.seg_width;
                const bool tr = p.transposed[seg] != 0;
                const bool stage_it = !tr || p.dup_out != nullptr;
                __half* ot = nullptr;
                if (tr) ot = reinterpret_cast<__half*>(p.out[seg]) + (static_cast<long long>(img) * p.seg_width + nloc0) * p.tok_pad + tok;
                for (int b = half; b < nblk; b += 2) {  // the two warps of a lane group take alternate 64-column blocks
                    const bool tail = b == p.epi_nfull;
                    const uint32_t rowp = smem_u32(stg) + b * 16384 + (tail ? r * 64 : r * 128);
                    const int sw = tail ? ((r >> 1) & 3) : (r & 7);
                    for (int cc = 0; cc < (tail ? 1 : 2); ++cc) {
                        const int c = 64 * b + 32 * cc;
                        uint32_t raw[32];
                        float v[32];
                        tmem_ld_32x32(t_row + c, raw);
                        if (GEGLU) {
                            uint32_t graw[32];
                            tmem_ld_32x32(t_row + bn_out + c, graw);
                            tmem_ld_wait();
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 b4 = lds128f(sb_a + (c + 4 * q) * 4);
                                const float4 g4 = lds128f(sb_a + (256 + c + 4 * q) * 4);
                                v[4 * q] = (__uint_as_float(raw[4 * q]) + b4.x) * gelu_erf_f(__uint_as_float(graw[4 * q]) + g4.x);
                                v[4 * q + 1] = (__uint_as_float(raw[4 * q + 1]) + b4.y) * gelu_erf_f(__uint_as_float(graw[4 * q + 1]) + g4.y);
                                v[4 * q + 2] = (__uint_as_float(raw[4 * q + 2]) + b4.z) * gelu_erf_f(__uint_as_float(graw[4 * q + 2]) + g4.z);
                                v[4 * q + 3] = (__uint_as_float(raw[4 * q + 3]) + b4.w) * gelu_erf_f(__uint_as_float(graw[4 * q + 3]) + g4.w);
                            }
                        } else {
                            tmem_ld_wait();
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const float4 b4 = lds128f(sb_a + (c + 4 * q) * 4);
                                v[4 * q] = __uint_as_float(raw[4 * q]) + b4.x;
                                v[4 * q + 1] = __uint_as_float(raw[4 * q + 1]) + b4.y;
                                v[4 * q + 2] = __uint_as_float(raw[4 * q + 2]) + b4.z;
                                v[4 * q + 3] = __uint_as_float(raw[4 * q + 3]) + b4.w;
                            }
                        }
                        if (p.rowbias && row_ok) {
                            const float* rb = p.rowbias + static_cast<long long>(img) * p.rowbias_ld + n0 + c;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + c + j < p.N) v[j] += __ldg(rb + j);
                        }
                        if (p.out_scale != 1.0f) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
                        }
                        if (tr && row_ok) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (n0 + c + j < p.N) ot[static_cast<long long>(c + j) * p.tok_pad] = __float2half_rn(v[j]);
                        }
                        if (stage_it)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t sp = rowp + (((4 * cc + q) ^ sw) << 4);
                            if (p.epi_res) {
                                const uint4 x = lds128(sp);
                                const __half2* h = reinterpret_cast<const __half2*>(&x);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 f = __half22float2(h[e]);
                                    v[q * 8 + e * 2] += f.x;
                                    v[q * 8 + e * 2 + 1] += f.y;
                                }
                            }
                            sts128(sp, pack_h2(v[q * 8 + 0], v[q * 8 + 1]), pack_h2(v[q * 8 + 2], v[q * 8 + 3]),
                                   pack_h2(v[q * 8 + 4], v[q * 8 + 5]), pack_h2(v[q * 8 + 6], v[q * 8 + 7]));
                        }
                    }
                }
                tc_fence_before();
                fence_proxy_async_smem();  // staging writes (generic proxy) -> visible to the bulk-store engine
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]);
                    const int r0 = lane_grp * 32;  // this warp's 32 rows are a rectangular sub-box of the tile's pixel box
                    const int cw = tw * p.bw + r0 % p.bw, ch = th * p.bh + (r0 / p.bw) % p.bh, cb = tb * p.nb + r0 / (p.bw * p.bh);
                    const uint32_t s0 = smem_u32(stg);
                    if (stage_it) {
                        const CUtensorMap* m64 = tr ? &em.dup[0] : &em.out[seg][0];
                        const CUtensorMap* m32 = tr ? &em.dup[1] : &em.out[seg][1];
                        for (int b = half; b < nblk; b += 2) {
                            if (b == p.epi_nfull) tma_store_4d(m32, s0 + b * 16384 + lane_grp * 2048, nloc0 + 64 * b, cw, ch, cb);
                            else tma_store_4d(m64, s0 + b * 16384 + lane_grp * 4096, nloc0 + 64 * b, cw, ch, cb);
                        }
                    }
                    bulk_commit_group();
                }
            } else if (p.splits == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {  // bn_out <= 256: at most four 32-column chunks per warp
                    const int c = 32 * half + 64 * i;
                    if (c >= bn_out) break;
                    uint32_t raw[32];
                    float v[32], g[GEGLU ? 32 : 1];
                    tmem_ld_32x32(t_row + c, raw);
                    if (GEGLU) {
                        uint32_t graw[32];
                        tmem_ld_32x32(t_row + bn_out + c, graw);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) g[j] = __uint_as_float(graw[j]);
                    } else {
                        tmem_ld_wait();
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                    epilogue_chunk<GEGLU>(p, v, g, c, bn_out, n0, row_ok, m, img, tok, sb);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_leader(&tempty[acc]); else mbar_arrive(&tempty[acc]); }
            } else {
                // ---- split-K: park this split's partial tile in its own fp32 workspace slice (plain stores)
                const int tile_mn = nt * m_tiles + mt;
                const long long slice = static_cast<long long>(GEMM_BM) * p.BN;
                // slice layout [BN / 4][128 rows][4 floats]: thread = row, so the 32 lanes of a warp touch 32 consecutive
                // 16-byte slots (512 contiguous bytes per instruction) both when parking and when reducing
                float* wrow0 = p.ws + static_cast<long long>(tile_mn) * p.splits * slice + static_cast<long long>(r) * 4;
                float* wrow = wrow0 + ks * slice;
                auto wofs = [](int col) { return static_cast<long long>(col >> 2) * (GEMM_BM * 4); };
                for (int c = 32 * half; c < p.BN; c += 64) {
                    uint32_t raw[32];
                    tmem_ld_32x32(t_row + c, raw);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (c + 4 * q < p.BN)
                            __stcg(reinterpret_cast<float4*>(wrow + wofs(c + 4 * q)),
                                   make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]),
                                               __uint_as_float(raw[4 * q + 2]), __uint_as_float(raw[4 * q + 3])));
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);  // the accumulator stage is free again
                __threadfence();
                asm volatile("bar.sync 1, %0;" ::"n"(32 * GEMM_EPI_WARPS) : "memory");
                if (warp == 2 && lane == 0) {
                    const unsigned int old = atomicAdd(&p.counters[tile_mn], 1u);
                    const int last = (old == static_cast<unsigned int>(p.splits - 1));
                    if (last) p.counters[tile_mn] = 0;  // self-cleaning: ready for the next launch
                    *last_flag = last;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * GEMM_EPI_WARPS) : "memory");
                const int is_last = *last_flag;
                if (is_last) {
                    __threadfence();
                    for (int c = 32 * half; c < bn_out; c += 64) {
                        float v[32], g[32];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = t4;
                            if (c + 4 * q < bn_out) {
                                for (int sl = 0; sl < p.splits; ++sl) {  // fixed order: deterministic sums
                                    const float4 x4 = __ldcg(reinterpret_cast<const float4*>(wrow0 + sl * slice + wofs(c + 4 * q)));
                                    t4.x += x4.x; t4.y += x4.y; t4.z += x4.z; t4.w += x4.w;
                                    if (GEGLU) {
                                        const float4 y4 = __ldcg(reinterpret_cast<const float4*>(wrow0 + sl * slice + wofs(bn_out + c + 4 * q)));
                                        g4.x += y4.x; g4.y += y4.y; g4.z += y4.z; g4.w += y4.w;
                                    }
                                }
                            }
                            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
                            g[4 * q] = g4.x; g[4 * q + 1] = g4.y; g[4 * q + 2] = g4.z; g[4 * q + 3] = g4.w;
                        }
                        epilogue_chunk<GEGLU>(p, v, g, c, bn_out, n0, row_ok, m, img, tok, sb);
                    }
                }
                // nobody may overwrite last_flag before every epilogue thread has read it
                asm volatile("bar.sync 1, %0;" ::"n"(32 * GEMM_EPI_WARPS) : "memory");
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (p.epi_tma && lane == 0) bulk_wait_group<0>();  // staging must outlive the bulk stores that read it
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();  // the peer may still be reading this CTA's shared memory / signalling its barriers
    if (warp == 2) {
        __syncwarp();
        if (PAIR) tmem_dealloc_2cta(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled get_tmap_encoder() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
    }
    return fn;
}

// fp16 tensor map with SWIZZLE_128B, zero OOB fill; dims innermost first; strides (bytes) for dims 1..rank-1.
static int make_tmap_f16_sw(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, CUtensorMapSwizzle swz);
int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
    return make_tmap_f16_sw(map, base, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B);
}
static int make_tmap_f16_sw(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                            const uint32_t* box, CUtensorMapSwizzle swz) {
    PFN_tmapEncodeTiled enc = get_tmap_encoder();
    if (!enc) return CTRLORA_ERR_TMAP;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "ctrlora: cuTensorMapEncodeTiled failed (%d) rank %d dims", (int)r, rank);
        for (int i = 0; i < rank; ++i) fprintf(stderr, " %llu", (unsigned long long)dims[i]);
        fprintf(stderr, " box");
        for (int i = 0; i < rank; ++i) fprintf(stderr, " %u", box[i]);
        fprintf(stderr, "\n");
        return CTRLORA_ERR_TMAP;
    }
    return CTRLORA_OK;
}

static int pow2_floor(int x) {
    int p = 1;
    while (p * 2 <= x) p *= 2;
    return p;
}

static int g_num_sms = 0;
static int g_sm_limit = 0;  // > 0: persistent grids leave SMs free for a concurrently running collective (ctrlora_set_sm_limit)
static bool g_attr_set = false;

}  // namespace ctrl

using namespace ctrl;

extern "C" int ctrlora_gemm_f16(const ctrlora_gemm_args* a, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !a->a || !a->w || !a->out[0]) return CTRLORA_ERR_ARG;
    if (a->a_c % 8 != 0 || a->a_ld % 8 != 0) return CTRLORA_ERR_ARG;
    if (a->kh != a->kw || (a->kh != 1 && a->kh != 3)) return CTRLORA_ERR_UNSUPPORTED;
    if (a->bf16) return CTRLORA_ERR_UNSUPPORTED;
    GemmKParams p;
    memset(&p, 0, sizeof(p));
    p.W = a->a_w; p.H = a->a_h; p.Bn = a->a_b;
    p.bw = pow2_floor(p.W < 128 ? p.W : 128);
    p.bh = pow2_floor(p.H < 128 / p.bw ? p.H : 128 / p.bw);
    p.nb = 128 / (p.bw * p.bh);
    p.tiles_w = (p.W + p.bw - 1) / p.bw;
    p.tiles_h = (p.H + p.bh - 1) / p.bh;
    p.tiles_b = (p.Bn + p.nb - 1) / p.nb;
    p.N = a->n;
    p.taps = a->kh * a->kw; p.kw = a->kw; p.pad = a->pad;
    p.kchunks = (a->a_c + GEMM_BK - 1) / GEMM_BK;
    p.kchunks2 = a->a2 ? (a->a2_c + GEMM_BK - 1) / GEMM_BK : 0;
    p.geglu = a->geglu;
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) return CTRLORA_ERR_CUDA;
    }
    const int m_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    const int k_iters = p.taps * p.kchunks + p.kchunks2;
    // ---- 2-CTA pairs (cta_group::2): every SM ingests half of the B tile; needs an even split of N and no split-K
    static int pair_env = -1;
    if (pair_env < 0) {
        const char* e = getenv("CTRLORA_GEMM_PAIR");
        pair_env = (e && e[0] == '0') ? 0 : (e && e[0] == '2') ? 2 : 1;  // 0: never, 1: per-shape policy, 2: wherever legal
    }
    // Measured (profiles/README.md): pairs win where the k loop dominates (3x3 convs and K >= 1024 linears: -10..-30 %),
    // lose on short-K 1x1 / GEGLU tiles whose time is the epilogue.
    const bool pair_shape = pair_env == 2 || (!p.geglu && k_iters >= (p.taps > 1 ? 24 : 16)) || (p.geglu && k_iters >= 20);
    const bool pair_ok = pair_env && pair_shape && g_num_sms >= 2 && a->force_single_cta == 0;
    // ---- pick the N tile and the K split with a per-tile cycle model, in cycles at the ~1.45 GHz the
    // part holds under tensor load: a k-step costs max(MMA = 2 x BN, operand bytes / 69 B/clk (~100 GB/s per SM, the
    // measured L2->SM share with all SMs pulling)); a launch costs waves x (k-steps + exposed epilogue).
    static int epi_env = -1;
    if (epi_env < 0) {
        const char* e = getenv("CTRLORA_GEMM_EPI");
        epi_env = (e && e[0] == '0') ? 0 : (e && e[0] == '2') ? 2 : 1;  // 0: direct only, 1: per-shape policy, 2: wherever legal
    }
    // TMA-staged epilogue (see below): legal for plain fp16 row-major outputs; wanted where the k loop is short
    bool epi_legal = !a->out_f32 && a->split_k <= 1 && a->ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(a->out[0]) & 15) == 0 &&
                     (!a->residual || (!a->residual_f32 && a->ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(a->residual) & 15) == 0));
    if (a->seg_width > 0) {  // q | k | v segments: plain segments need aligned bases, a transposed one may carry a duplicate
        const int nseg = (a->n + a->seg_width - 1) / a->seg_width;
        if (a->residual || nseg > 3 || a->seg_width % 8 != 0) epi_legal = false;
        for (int i = 0; i < nseg && i < 3; ++i)
            if (!a->transposed[i] && (!a->out[i] || (reinterpret_cast<uintptr_t>(a->out[i]) & 15) != 0)) epi_legal = false;
        if (a->dup_out && ((reinterpret_cast<uintptr_t>(a->dup_out) & 15) != 0 || a->dup_ld % 8 != 0)) epi_legal = false;
    } else if (a->transposed[0] || a->dup_out) {
        epi_legal = false;
    }
    static int epi_kmax = -1;
    if (epi_kmax < 0) {
        const char* e = getenv("CTRLORA_GEMM_EPI_KMAX");
        epi_kmax = e ? atoi(e) : 32;
    }
    const bool epi_wanted = epi_legal && (epi_env == 2 || (epi_env == 1 && k_iters <= epi_kmax));
    int bn_out = a->block_n, splits = a->split_k > 0 ? a->split_k : 1;
    if (bn_out <= 0) {
        double best_cost = -1;
        const int max_out = p.geglu ? 128 : 256;
        static const int split_cands[] = {1, 2, 3, 4, 6, 8, 12, 16};
        for (int cand = max_out; cand >= 16; cand -= 16) {
            if (a->seg_width > 0 && a->seg_width % cand != 0) continue;
            if (epi_wanted && cand % 32 != 0) continue;  // the staged epilogue works in 64- and 32-column blocks
            const int bnt = p.geglu ? 2 * cand : cand;
            const int nt = (p.N + cand - 1) / cand;
            const long tiles_mn = (long)m_tiles * nt;
            const double waste = (double)nt * cand / p.N;  // columns computed beyond N
            for (int si = 0; si < 8; ++si) {
                int S = split_cands[si];
                if (a->split_k > 0) { if (si > 0) break; S = a->split_k; }
                if (S > 1) {
                    if (a->split_k <= 0) break;  // measured (tools/sweep_gemm.py): the atomic-add reduction costs more than
                                                 // it saves at every shape of the 512x512 path; only an explicit request splits
                    if (!a->splitk_ws || !a->splitk_counters) break;
                    if ((long long)tiles_mn * S * GEMM_BM * bnt * 4 > a->splitk_ws_bytes || tiles_mn > a->splitk_counters_len) break;
                    if (a->seg_width > 0 && a->transposed[0] + a->transposed[1] + a->transposed[2] > 0 && false) break;
                }
                const int kps = (k_iters + S - 1) / S;
                if (S > 1 && kps < 4 && a->split_k <= 0) break;
                const int s_eff = (k_iters + kps - 1) / kps;
                const bool cpair = pair_ok && s_eff == 1 && bnt % 32 == 0;
                const long tiles = cpair ? (long)((m_tiles + 1) / 2) * nt : tiles_mn * s_eff;
                const long slots = cpair ? g_num_sms / 2 : g_num_sms;
                const long waves = (tiles + slots - 1) / slots;
                const double t_mma = kps * 4.0 * (bnt / 2 > 32 ? bnt / 2 : 32);
                const double t_load = kps * (double)(GEMM_A_BYTES + (cpair ? bnt / 2 : bnt) * 128) / 69.0;
                const double t_epi = (cand / 32 + 1) * 350.0;
                double t_tile = (t_mma > t_load ? t_mma : t_load);
                if (t_epi > t_tile) t_tile = t_epi;
                t_tile += 600.0;
                if (s_eff > 1) t_tile += bnt * 12.0 + t_epi;
                const double cost = waves * t_tile * (0.5 + 0.5 * waste);
                if (best_cost < 0 || cost < best_cost) { best_cost = cost; bn_out = cand; splits = s_eff; }
            }
        }
    }
    if (bn_out % 16 != 0 || bn_out < 16 || bn_out > (p.geglu ? 128 : 256)) return CTRLORA_ERR_ARG;
    if (a->seg_width > 0 && a->seg_width % bn_out != 0) return CTRLORA_ERR_ARG;
    p.BN = p.geglu ? 2 * bn_out : bn_out;
    p.n_tiles = (p.N + bn_out - 1) / bn_out;
    p.kiters_per_split = (k_iters + splits - 1) / splits;
    p.splits = (k_iters + p.kiters_per_split - 1) / p.kiters_per_split;
    if (p.splits > 1) {
        const long long tiles_mn = (long long)m_tiles * p.n_tiles;
        if (!a->splitk_ws || !a->splitk_counters || tiles_mn * p.splits * GEMM_BM * p.BN * 4 > a->splitk_ws_bytes ||
            tiles_mn > a->splitk_counters_len)
            return CTRLORA_ERR_ARG;
        p.ws = a->splitk_ws;
        p.counters = a->splitk_counters;
    }
    const bool pair = pair_ok && p.splits == 1 && (p.BN % 32 == 0);
    const int b_rows = pair ? p.BN / 2 : p.BN;  // B rows held by one CTA
    p.stage_bytes = GEMM_A_BYTES + ((b_rows * 128 + 1023) / 1024) * 1024;
    // ---- epilogue flavour: short-K shapes are epilogue-bound (profiles/README.md: per-row 16-byte global accesses of the
    // direct epilogue cost ~8 us per 128x160 tile), so they stage the tile in swizzled shared memory and let TMA do the
    // (coalesced, asynchronous) residual load and output store. Long-K shapes keep the deeper operand ring instead.
    int ring_bytes = GEMM_SMEM_DATA;
    {
        if (epi_wanted && p.splits == 1 && bn_out % 32 == 0) {
            const int nfull = bn_out / 64, tail = (bn_out % 64) ? 1 : 0;
            const int buf_bytes = nfull * 16384 + tail * 8192;
            int nbuf = 2;
            if ((GEMM_SMEM_DATA - 2 * buf_bytes) / p.stage_bytes < 3) nbuf = 1;
            static int nbuf_env = -1;  // CTRLORA_GEMM_EPI_NBUF=1: single staging buffer -> one or two more ring stages in flight
            if (nbuf_env < 0) {
                const char* e = getenv("CTRLORA_GEMM_EPI_NBUF");
                nbuf_env = e ? atoi(e) : 0;
            }
            if (nbuf_env == 1) nbuf = 1;
            if ((GEMM_SMEM_DATA - nbuf * buf_bytes) / p.stage_bytes >= 3) {
                p.epi_tma = 1; p.epi_nbuf = nbuf; p.epi_res = a->residual ? 1 : 0;
                p.epi_nfull = nfull; p.epi_tail = tail; p.epi_buf_bytes = buf_bytes;
                p.epi_off = GEMM_SMEM_DATA - nbuf * buf_bytes;
                ring_bytes = p.epi_off;
            }
        }
    }
    p.stages = ring_bytes / p.stage_bytes;
    if (p.stages > GEMM_MAX_STAGES) p.stages = GEMM_MAX_STAGES;
    p.idesc = umma_idesc_f16(pair ? 2 * GEMM_BM : GEMM_BM, p.BN, 0);
    for (int i = 0; i < 3; ++i) { p.out[i] = a->out[i]; p.transposed[i] = a->transposed[i]; }
    p.seg_width = a->seg_width;
    p.ldc = a->ldc; p.out_f32 = a->out_f32;
    p.bias = a->bias; p.rowbias = a->rowbias;
    p.rows_per_img = a->rows_per_img > 0 ? a->rows_per_img : p.W * p.H;
    p.residual = reinterpret_cast<const __half*>(a->residual); p.ldr = a->ldr;
    p.residual_f32 = a->residual_f32;
    p.rowbias_ld = a->rowbias_ld > 0 ? a->rowbias_ld : a->n;
    p.out_scale = a->out_scale;
    p.head_dim = a->head_dim; p.tok_pad = a->tok_pad;
    p.dup_out = reinterpret_cast<__half*>(a->dup_out); p.dup_ld = a->dup_ld;

    CUtensorMap tmA, tmB, tmA2, tmB2;
    {
        uint64_t dims[4] = {(uint64_t)a->a_c, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.Bn};
        uint64_t str[3] = {(uint64_t)a->a_ld * 2, (uint64_t)a->a_ld * 2 * p.W, (uint64_t)a->a_ld * 2 * p.W * p.H};
        uint32_t box[4] = {GEMM_BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.nb};
        int rc = make_tmap_f16(&tmA, a->a, 4, dims, str, box);
        if (rc) return rc;
        const uint64_t rows = p.geglu ? 2ull * p.N : (uint64_t)p.N;
        uint64_t wd[3] = {(uint64_t)a->a_c, (uint64_t)p.taps, rows};
        uint64_t ws[2] = {(uint64_t)a->a_c * 2, (uint64_t)a->a_c * 2 * p.taps};
        uint32_t wb[3] = {GEMM_BK, 1, (uint32_t)((pair && !p.geglu) ? p.BN / 2 : bn_out)};
        rc = make_tmap_f16(&tmB, a->w, 3, wd, ws, wb);
        if (rc) return rc;
    }
    if (a->a2) {
        if (!a->w2 || a->a2_c % 8 != 0 || a->a2_ld % 8 != 0 || p.geglu) return CTRLORA_ERR_ARG;
        uint64_t dims[4] = {(uint64_t)a->a2_c, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.Bn};
        uint64_t str[3] = {(uint64_t)a->a2_ld * 2, (uint64_t)a->a2_ld * 2 * p.W, (uint64_t)a->a2_ld * 2 * p.W * p.H};
        uint32_t box[4] = {GEMM_BK, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.nb};
        int rc = make_tmap_f16(&tmA2, a->a2, 4, dims, str, box);
        if (rc) return rc;
        uint64_t wd[3] = {(uint64_t)a->a2_c, 1, (uint64_t)p.N};
        uint64_t ws[2] = {(uint64_t)a->a2_c * 2, (uint64_t)a->a2_c * 2};
        uint32_t wb[3] = {GEMM_BK, 1, (uint32_t)(pair ? p.BN / 2 : bn_out)};
        rc = make_tmap_f16(&tmB2, a->w2, 3, wd, ws, wb);
        if (rc) return rc;
    } else {
        tmA2 = tmA;
        tmB2 = tmB;
    }
    GemmEpiMaps em;
    for (int i = 0; i < 3; ++i) { em.out[i][0] = tmA; em.out[i][1] = tmA; }
    em.dup[0] = em.dup[1] = em.res[0] = em.res[1] = tmA;
    if (p.epi_tma) {
        // output: per-warp stores of a 32-row sub-box of the tile's (bw, bh, nb) pixel box; residual: whole-tile loads
        const int sw = p.bw < 32 ? p.bw : 32;
        const int sh = p.bh < 32 / sw ? p.bh : 32 / sw;
        const int sb = 32 / (sw * sh);
        auto pair_of = [&](CUtensorMap* dst, const void* base, long long ld, int cols, int bw_, int bh_, int nb_) -> int {
            uint64_t dims[4] = {(uint64_t)cols, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.Bn};
            uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * p.W, (uint64_t)ld * 2 * p.W * p.H};
            uint32_t box[4] = {64, (uint32_t)bw_, (uint32_t)bh_, (uint32_t)nb_};
            int rc = make_tmap_f16_sw(&dst[0], base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
            if (rc) return rc;
            box[0] = 32;
            return make_tmap_f16_sw(&dst[1], base, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
        };
        if (a->seg_width > 0) {
            const int nseg = (p.N + a->seg_width - 1) / a->seg_width;
            for (int i = 0; i < nseg; ++i) {
                if (a->transposed[i]) continue;
                int rc = pair_of(em.out[i], a->out[i], a->ldc, a->seg_width, sw, sh, sb);
                if (rc) return rc;
            }
            if (a->dup_out) {
                int rc = pair_of(em.dup, a->dup_out, a->dup_ld, a->seg_width, sw, sh, sb);
                if (rc) return rc;
            }
        } else {
            int rc = pair_of(em.out[0], a->out[0], a->ldc, p.N, sw, sh, sb);
            if (rc) return rc;
        }
        if (p.epi_res) {
            int rc = pair_of(em.res, a->residual, a->ldr, p.N, p.bw, p.bh, p.nb);
            if (rc) return rc;
        }
    }
    if (!g_attr_set) {
        if (cudaFuncSetAttribute(gemm_tcgen05_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_tcgen05_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_tcgen05_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_tcgen05_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES) != cudaSuccess)
            return CTRLORA_ERR_CUDA;
        g_attr_set = true;
    }
    cudaError_t lrc;
    if (pair) {
        const int units = ((m_tiles + 1) / 2) * p.n_tiles;           // one unit = one 256-row tile for one CTA pair
        int clusters = (g_sm_limit > 0 && g_sm_limit < g_num_sms ? g_sm_limit : g_num_sms) / 2;
        if (units < clusters) clusters = units;
        const dim3 grid2(2 * clusters), block(GEMM_THREADS);
        lrc = p.geglu ? launch_cluster(gemm_tcgen05_kernel<true, true>, grid2, block, (size_t)GEMM_SMEM_BYTES, stream, 2u, tmA, tmB, tmA2, tmB2, em, p)
                      : launch_cluster(gemm_tcgen05_kernel<false, true>, grid2, block, (size_t)GEMM_SMEM_BYTES, stream, 2u, tmA, tmB, tmA2, tmB2, em, p);
    } else {
        const int total = m_tiles * p.n_tiles * p.splits;
        const int sms_avail = g_sm_limit > 0 && g_sm_limit < g_num_sms ? g_sm_limit : g_num_sms;
        const int grid = total < sms_avail ? total : sms_avail;
        lrc = p.geglu ? launch_pdl(gemm_tcgen05_kernel<true, false>, dim3(grid), dim3(GEMM_THREADS), (size_t)GEMM_SMEM_BYTES,
                                   stream, tmA, tmB, tmA2, tmB2, em, p)
                      : launch_pdl(gemm_tcgen05_kernel<false, false>, dim3(grid), dim3(GEMM_THREADS), (size_t)GEMM_SMEM_BYTES,
                                   stream, tmA, tmB, tmA2, tmB2, em, p);
    }
    if (lrc != cudaSuccess) return CTRLORA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

// Persistent GEMM grids use at most `limit` SMs (0 = all).  A communication kernel that runs next to the backward (the
// overlapped gradient all-reduce) owns a few SMs; a 148-CTA persistent grid would otherwise wait for them and run a second,
// nearly empty wave.  The limit is read at launch time, i.e. it is baked into a CUDA graph at capture.
extern "C" int ctrlora_set_sm_limit(int limit) {
    if (limit < 0) return CTRLORA_ERR_ARG;
    g_sm_limit = limit;
    return CTRLORA_OK;
}
