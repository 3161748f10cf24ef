// Attention backward on tcgen05 (training): given Q, K, V, O, dO and the forward's log-sum-exp, produce dQ, dK, dV.
// reference: autograd through CrossAttention.forward, ldm/modules/attention.py:163-194.
//
// Two kernels, no atomics (fp32 global atomics were measured at ~40 G/s on this part, far too slow for dQ):
//   attn_bwd_dq_kernel    CTA = 128 queries of one (image, head); loops over key tiles:
//                           S = Q K^T, dP = dO V^T (UMMA) -> P = exp2(S c - lse), dS = P (dP - D) (row threads)
//                           -> dQ += dS K (UMMA, K tile re-read as an MN-major B operand)
//   attn_bwd_dkdv_kernel  CTA = 128 keys of one (image, head); loops over query tiles:
//                           S^T = K Q^T, dP^T = V dO^T (UMMA) -> P^T, dS^T (thread = key row, lse/D per column)
//                           -> dV += P^T dO, dK += dS^T Q (UMMA, the Q / dO tiles re-read as MN-major B operands)
// The same [rows][64-col] SWIZZLE_128B TMA tiles serve both as K-major operands (contraction over d) and as MN-major
// operands (contraction over tokens); only the UMMA descriptor differs.  D = rowsum(dO * O) comes from a small pre-pass.
#include "common.cuh"
#include "ctrlora_b200.h"
#include "gemm_sm100.cuh"
#include <stdlib.h>

#ifdef CTRLORA_SPIN_WAIT
#define MBAR_CHAIN_WAIT(bar, par) mbar_wait_spin(bar, par)
#else
#define MBAR_CHAIN_WAIT(bar, par) mbar_wait(bar, par)
#endif

namespace ctrl {
#ifdef CTRLORA_TIMELINE
__device__ long long g_tl[4096];
#define TL(slot) do { if (blockIdx.x == 3 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0 && (slot) < 4096) g_tl[slot] = clock64(); } while (0)
#else
#define TL(slot) do { } while (0)
#endif

int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

struct AttnBwdParams {
    int Nq, Nk, heads, d, d16, nkc;
    float scale, scale_log2e;
    const float* lse;    // [B, H, Nq]  log2-domain: P = exp2(s * scale_log2e - lse)
    const float* delta;  // [B, H, Nq]  rowsum(dO * O)
    __half* dq; long long lddq;
    __half* dk; long long lddk;
    __half* dv; long long lddv;
    uint32_t idesc_s;    // M=128, N = tile width (keys for dq kernel, queries for dkdv kernel), K-major operands
    uint32_t idesc_acc;  // M=128, N = d16, A K-major, B MN-major
};

constexpr int AB_THREADS = 192;  // warps 0-3: row threads, warp 4: TMA, warp 5: MMA

// write 32 fp16 values (packed pairs) of row r, columns [c, c+32) of a K-major SWIZZLE_128B operand tile made of
// 64-column chunks of `rows` rows each
__device__ __forceinline__ void store_row_chunk(uint32_t tile, int rows, int r, int c, const uint32_t* packed) {
    const uint32_t chunk = tile + (c >> 6) * rows * 128 + r * 128;
    const int u0 = (c & 63) >> 3;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        sts128(chunk + (((u0 + u) ^ (r & 7)) << 4), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
}

// ================================================================================================ D = rowsum(dO * O)
__global__ void attn_bwd_delta_kernel(const __half* __restrict__ o, long long ldo, const __half* __restrict__ dout, long long lddo,
                                      float* __restrict__ delta, int batch, int heads, int nq, int d) {
    pdl_launch_dependents();
    pdl_wait();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;  // (b, q, h)
    if (i >= static_cast<long long>(batch) * nq * heads) return;
    const int h = static_cast<int>(i % heads);
    const long long row = i / heads;  // b * nq + q
    const int b = static_cast<int>(row / nq), q = static_cast<int>(row % nq);
    const __half* op = o + row * ldo + h * d;
    const __half* dp = dout + row * lddo + h * d;
    float acc = 0.f;
    for (int c = 0; c < d; c += 8) {
        uint4 u = *reinterpret_cast<const uint4*>(op + c), w = *reinterpret_cast<const uint4*>(dp + c);
        const __half2* a = reinterpret_cast<const __half2*>(&u);
        const __half2* bb = reinterpret_cast<const __half2*>(&w);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 x = __half22float2(a[e]), y = __half22float2(bb[e]);
            acc += x.x * y.x + x.y * y.y;
        }
    }
    delta[(static_cast<long long>(b) * heads + h) * nq + q] = acc;
}

// ================================================================================================ dQ
template <int DPAD, int BKV>
struct DqSmem {
    static constexpr int NKC = (DPAD + 63) / 64;
    static constexpr int Q_BYTES = NKC * 128 * 128;      // Q and dO: [nkc][128 q][128 B]
    static constexpr int KV_BYTES = NKC * BKV * 128;     // K and V:  [nkc][BKV keys][128 B]
    static constexpr int DS_BYTES = (BKV / 64) * 128 * 128 > 0 ? (BKV / 64) * 128 * 128 : 128 * 128;
    // K/V tiles are double-buffered where shared memory allows: the next tile's S / dP products are then issued before
    // the current tile's dQ accumulation and the TMA latency disappears from the per-tile chain
    static constexpr int STAGES = (2 * Q_BYTES + 4 * KV_BYTES + DS_BYTES <= 200 * 1024) ? 2 : 1;
    static constexpr int STAGE_BYTES = 2 * KV_BYTES;
    static constexpr int OFF_DO = Q_BYTES, OFF_K = 2 * Q_BYTES, OFF_V = OFF_K + KV_BYTES, OFF_DS = OFF_K + STAGES * STAGE_BYTES;
    static constexpr int DATA = OFF_DS + DS_BYTES;
    static constexpr int TOTAL = DATA + 1024 + 128;
};

// TS: every A operand of the dQ kernel lives in TMEM (d <= 48, 64-key tiles: S 64 | dP 64 | dQ 48 | dS 32 | Q 24 | dO 24
// = 256 columns). Q and dO rows are copied once from their TMA tiles by the thread that owns the row; dS is written
// straight from the row math. An A-from-TMEM MMA costs N/2 cycles per k-step instead of 32 + N/4 (the smem A read,
// tools/microbench/mma_issue.cu): 288 instead of 464 tensor cycles per 128x64 tile.
template <int DPAD, int BKV>
struct DqTs { static constexpr bool ON = (DPAD == 48 && BKV == 64); };

template <int DPAD, int BKV>
__global__ void __launch_bounds__(AB_THREADS, (2 * BKV + DPAD <= 256) ? 2 : 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const __grid_constant__ AttnBwdParams p) {
    using L = DqSmem<DPAD, BKV>;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem, *sDO = smem + L::OFF_DO, *sDS = smem + L::OFF_DS;
    constexpr int ST = L::STAGES;
    auto sK = [&](int st) { return smem + L::OFF_K + st * L::STAGE_BYTES; };
    auto sV = [&](int st) { return smem + L::OFF_V + st * L::STAGE_BYTES; };
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::DATA);
    uint64_t *q_full = bars, *kv_full = bars + 1 /*[2]*/, *kv_free = bars + 3 /*[2]*/, *s_full = bars + 5, *ds_full = bars + 6,
             *acc_done = bars + 7, *a_ready = bars + 8;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 9);
    constexpr bool TS = DqTs<DPAD, BKV>::ON;
    constexpr uint32_t TMEM_COLS = (2 * BKV + DPAD <= 256) ? 256 : 512;  // 256 columns: two CTAs share an SM's TMEM
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
    if (warp == 4 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    if (warp == 5 && lane == 0) {
        mbar_init(q_full, 1); mbar_init(s_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_free[i], 1); }
        mbar_init(ds_full, 128); mbar_init(acc_done, 1); mbar_init(a_ready, 128);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    const uint32_t tm_s = tmem_base, tm_dp = tmem_base + BKV, tm_dq = tmem_base + 2 * BKV;
    const uint32_t tm_ds16 = tm_dq + DPAD, tm_q16 = tm_ds16 + BKV / 2, tm_do16 = tm_q16 + DPAD / 2;  // TS only
    const int n_tiles = (p.Nk + BKV - 1) / BKV;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * p.nkc * 128 * 128);
            for (int kc = 0; kc < p.nkc; ++kc) {
                tma_load_4d(sQ + kc * 128 * 128, &tmQ, q_full, kc * 64, head, q0, img);
                tma_load_4d(sDO + kc * 128 * 128, &tmDO, q_full, kc * 64, head, q0, img);
            }
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % ST;
                if (j >= ST) mbar_wait(&kv_free[st], ((j / ST) - 1) & 1);
                mbar_expect_tx(&kv_full[st], 2 * p.nkc * BKV * 128);
                for (int kc = 0; kc < p.nkc; ++kc) {
                    tma_load_4d(sK(st) + kc * BKV * 128, &tmK, &kv_full[st], kc * 64, head, j * BKV, img);
                    tma_load_4d(sV(st) + kc * BKV * 128, &tmV, &kv_full[st], kc * 64, head, j * BKV, img);
                }
            }
        }
    } else if (warp == 5) {
        // whole warp in the loop, one elected lane issues (uniform registers for the tcgen05 operands)
        // descriptors are built once; per-MMA operands differ by compile-time offsets in the 16-byte address field
        // (measured with the clock timeline: the first version needed ~650-750 cycles to ISSUE one tile's MMAs, which sat
        // on the per-tile chain twice: before s_full and before kv_free)
        constexpr int KS = DPAD / 16;  // k-steps over d (columns >= d are TMA zero fill)
        const uint64_t dQ = umma_desc_kmajor_sw128(smem_u32(sQ)), dDO = umma_desc_kmajor_sw128(smem_u32(sDO)),
                       dDS = umma_desc_kmajor_sw128(smem_u32(sDS));
        const uint32_t idesc_s = p.idesc_s, idesc_acc = p.idesc_acc;
        auto issue_s_dp = [&](int j) {  // S = Q K(j)^T, dP = dO V(j)^T
            const int st = j % ST;
            mbar_wait(&kv_full[st], (j / ST) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(sK(st))), dV = umma_desc_kmajor_sw128(smem_u32(sV(st)));
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t oq = ((ks >> 2) * 128 * 128 + (ks & 3) * 32) >> 4, ok = ((ks >> 2) * BKV * 128 + (ks & 3) * 32) >> 4;
                    if (TS) umma_f16_ts(tm_s, tm_q16 + 8 * ks, dK + ok, idesc_s, ks ? 1u : 0u);
                    else umma_f16(tm_s, dQ + oq, dK + ok, idesc_s, ks ? 1u : 0u);
                }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t oq = ((ks >> 2) * 128 * 128 + (ks & 3) * 32) >> 4, ok = ((ks >> 2) * BKV * 128 + (ks & 3) * 32) >> 4;
                    if (TS) umma_f16_ts(tm_dp, tm_do16 + 8 * ks, dV + ok, idesc_s, ks ? 1u : 0u);
                    else umma_f16(tm_dp, dDO + oq, dV + ok, idesc_s, ks ? 1u : 0u);
                }
                umma_commit(s_full);
            }
            __syncwarp();
        };
        if (TS) mbar_wait(a_ready, 0);  // the row threads have copied their Q / dO rows into TMEM
        else mbar_wait(q_full, 0);
        issue_s_dp(0);
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % ST;
            MBAR_CHAIN_WAIT(ds_full, j & 1);  // dS(j) is in shared memory; S / dP(j) have been read out
            if (ST == 2 && j + 1 < n_tiles) issue_s_dp(j + 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dKm = desc_mn_sw128(smem_u32(sK(st)), BKV * 128);
                const uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {  // contraction over the keys of this tile
                    const uint32_t oa = ((ks >> 2) * 128 * 128 + (ks & 3) * 32) >> 4;
                    if (TS) umma_f16_ts(tm_dq, tm_ds16 + 8 * ks, dKm + ((ks * 2048) >> 4), idesc_acc, (acc | ks) ? 1u : 0u);
                    else umma_f16(tm_dq, dDS + oa, dKm + ((ks * 2048) >> 4), idesc_acc, (acc | ks) ? 1u : 0u);
                }
                umma_commit(&kv_free[st]);  // K/V stage and the dS tile are free once these MMAs complete
                if (j + 1 == n_tiles) umma_commit(acc_done);
            }
            __syncwarp();
            if (ST == 1 && j + 1 < n_tiles) issue_s_dp(j + 1);
        }
    } else {
        const int r = warp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const bool row_ok = q0 + r < p.Nq;
        const long long stat_idx = (static_cast<long long>(img) * p.heads + head) * p.Nq + q0 + r;
        const float lse = row_ok ? p.lse[stat_idx] : 0.f;
        const float dl = row_ok ? p.delta[stat_idx] : 0.f;
        if (TS) {
            // row r of the Q / dO TMA tiles (SWIZZLE_128B: 16-byte unit u sits at u ^ (r & 7); columns >= d are zero fill)
            mbar_wait(q_full, 0);
            const uint32_t rq = smem_u32(sQ) + r * 128, rdo = smem_u32(sDO) + r * 128;
            uint32_t w[24];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const uint4 x = lds128(rq + ((u ^ (r & 7)) << 4));
                w[4 * u] = x.x; w[4 * u + 1] = x.y; w[4 * u + 2] = x.z; w[4 * u + 3] = x.w;
            }
            tmem_st_32x16(tm_q16 + lane_off, w);
            tmem_st_32x8(tm_q16 + lane_off + 16, w + 16);
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const uint4 x = lds128(rdo + ((u ^ (r & 7)) << 4));
                w[4 * u] = x.x; w[4 * u + 1] = x.y; w[4 * u + 2] = x.z; w[4 * u + 3] = x.w;
            }
            tmem_st_32x16(tm_do16 + lane_off, w);
            tmem_st_32x8(tm_do16 + lane_off + 16, w + 16);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(a_ready);
        }
        for (int j = 0; j < n_tiles; ++j) {
            MBAR_CHAIN_WAIT(s_full, j & 1);
            tc_fence_after();
            const int kv_valid = min(BKV, p.Nk - j * BKV);
#pragma unroll 1
            for (int c = 0; c < BKV; c += 32) {
                uint32_t sr[32], dr[32], packed[16];
                tmem_ld_32x32(tm_s + lane_off + c, sr);
                tmem_ld_32x32(tm_dp + lane_off + c, dr);
                tmem_ld_wait();
                // the dQ MMAs of the previous tile read the dS buffer: they must be done before it is overwritten
                if (c == 0 && j > 0) mbar_wait(&kv_free[(j - 1) % ST], ((j - 1) / ST) & 1);
                if (kv_valid == BKV) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float p0 = fast_exp2(fmaf(__uint_as_float(sr[i]), p.scale_log2e, -lse));
                        const float p1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), p.scale_log2e, -lse));
                        packed[i >> 1] = pack_half2(p0 * (__uint_as_float(dr[i]) - dl), p1 * (__uint_as_float(dr[i + 1]) - dl));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float ds0 = 0.f, ds1 = 0.f;
                        if (c + i < kv_valid) {
                            const float pr = fast_exp2(fmaf(__uint_as_float(sr[i]), p.scale_log2e, -lse));
                            ds0 = pr * (__uint_as_float(dr[i]) - dl);
                        }
                        if (c + i + 1 < kv_valid) {
                            const float pr = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), p.scale_log2e, -lse));
                            ds1 = pr * (__uint_as_float(dr[i + 1]) - dl);
                        }
                        packed[i >> 1] = pack_half2(ds0, ds1);
                    }
                }
                if (TS) tmem_st_32x16(tm_ds16 + lane_off + (c >> 1), packed);
                else store_row_chunk(smem_u32(sDS), 128, r, c, packed);
            }
            if (TS) tmem_st_wait();
            else fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(ds_full);
        }
        mbar_wait(acc_done, 0);
        tc_fence_after();
        __half* out = p.dq + (static_cast<long long>(img) * p.Nq + q0 + r) * p.lddq + head * p.d;
#pragma unroll
        for (int c = 0; c < DPAD; c += 16) {
            if (c < p.d) {
                uint32_t raw[16];
                tmem_ld_32x16(tm_dq + lane_off + c, raw);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (c + g * 8 < p.d) {
                            uint4 u;
                            uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                w[e] = pack_half2(__uint_as_float(raw[g * 8 + 2 * e]) * p.scale, __uint_as_float(raw[g * 8 + 2 * e + 1]) * p.scale);
                            *reinterpret_cast<uint4*>(out + c + g * 8) = u;
                        }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ================================================================================================ dK, dV
template <int DPAD, int BQ>
struct DkvSmem {
    static constexpr int NKC = (DPAD + 63) / 64;
    static constexpr int KV_BYTES = NKC * 128 * 128;   // K and V: [nkc][128 keys][128 B]
    static constexpr int Q_BYTES = NKC * BQ * 128;     // Q and dO: [nkc][BQ queries][128 B]
    static constexpr int PT_BYTES = (BQ / 64) * 128 * 128;  // P^T and dS^T: [BQ/64][128 keys][128 B]
    // Q / dO tiles are double-buffered where shared memory allows (see DqSmem)
    static constexpr int STAGES = (2 * KV_BYTES + 4 * Q_BYTES + 2 * PT_BYTES <= 200 * 1024) ? 2 : 1;
    static constexpr int STAGE_BYTES = 2 * Q_BYTES;
    static constexpr int OFF_V = KV_BYTES, OFF_Q = 2 * KV_BYTES, OFF_DO = OFF_Q + Q_BYTES, OFF_PT = OFF_Q + STAGES * STAGE_BYTES,
                         OFF_DST = OFF_PT + PT_BYTES, OFF_STAT = OFF_DST + PT_BYTES;
    static constexpr int DATA = OFF_STAT + 2 * 2 * BQ * 4;  // lse / delta of the query tile, double-buffered
    static constexpr int BAR = (DATA + 127) / 128 * 128;
    static constexpr int TOTAL = BAR + 1024 + 128;
};

template <int DPAD, int BQ>
__global__ void __launch_bounds__(AB_THREADS, (2 * BQ + 2 * DPAD <= 256) ? 2 : 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                     const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ AttnBwdParams p) {
    using L = DkvSmem<DPAD, BQ>;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sK = smem, *sV = smem + L::OFF_V, *sPT = smem + L::OFF_PT, *sDST = smem + L::OFF_DST;
    constexpr int ST = L::STAGES;
    auto sQ = [&](int st) { return smem + L::OFF_Q + st * L::STAGE_BYTES; };
    auto sDO = [&](int st) { return smem + L::OFF_DO + st * L::STAGE_BYTES; };
    float* sStat = reinterpret_cast<float*>(smem + L::OFF_STAT);  // [2 buffers][lse BQ | delta BQ]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
    uint64_t *kv_full = bars, *q_full = bars + 1 /*[2]*/, *q_free = bars + 3 /*[2]*/, *s_full = bars + 5, *p_full = bars + 6,
             *acc_done = bars + 7;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
    constexpr uint32_t TMEM_COLS = (2 * BQ + 2 * DPAD <= 256) ? 256 : 512;
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
    if (warp == 4 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    if (warp == 5 && lane == 0) {
        mbar_init(kv_full, 1); mbar_init(s_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_free[i], 1); }
        mbar_init(p_full, 128); mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    const uint32_t tm_s = tmem_base, tm_dp = tmem_base + BQ, tm_dv = tmem_base + 2 * BQ, tm_dk = tm_dv + DPAD;
    const int n_tiles = (p.Nq + BQ - 1) / BQ;

    if (warp == 4) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * p.nkc * 128 * 128);
            for (int kc = 0; kc < p.nkc; ++kc) {
                tma_load_4d(sK + kc * 128 * 128, &tmK, kv_full, kc * 64, head, k0, img);
                tma_load_4d(sV + kc * 128 * 128, &tmV, kv_full, kc * 64, head, k0, img);
            }
            for (int i = 0; i < n_tiles; ++i) {
                const int st = i % ST;
                if (i >= ST) mbar_wait(&q_free[st], ((i / ST) - 1) & 1);
                mbar_expect_tx(&q_full[st], 2 * p.nkc * BQ * 128);
                for (int kc = 0; kc < p.nkc; ++kc) {
                    tma_load_4d(sQ(st) + kc * BQ * 128, &tmQ, &q_full[st], kc * 64, head, i * BQ, img);
                    tma_load_4d(sDO(st) + kc * BQ * 128, &tmDO, &q_full[st], kc * 64, head, i * BQ, img);
                }
            }
        }
    } else if (warp == 5) {
        constexpr int KS = DPAD / 16;  // k-steps over d (columns >= d are TMA zero fill)
        const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(sK)), dV = umma_desc_kmajor_sw128(smem_u32(sV)),
                       dPT = umma_desc_kmajor_sw128(smem_u32(sPT)), dDST = umma_desc_kmajor_sw128(smem_u32(sDST));
        const uint32_t idesc_s = p.idesc_s, idesc_acc = p.idesc_acc;
        auto issue_s_dp = [&](int i) {  // S^T = K Q(i)^T, dP^T = V dO(i)^T
            const int st = i % ST;
            mbar_wait(&q_full[st], (i / ST) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dQ = umma_desc_kmajor_sw128(smem_u32(sQ(st))), dDO = umma_desc_kmajor_sw128(smem_u32(sDO(st)));
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t ok = ((ks >> 2) * 128 * 128 + (ks & 3) * 32) >> 4, oq = ((ks >> 2) * BQ * 128 + (ks & 3) * 32) >> 4;
                    umma_f16(tm_s, dK + ok, dQ + oq, idesc_s, ks ? 1u : 0u);
                }
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t ok = ((ks >> 2) * 128 * 128 + (ks & 3) * 32) >> 4, oq = ((ks >> 2) * BQ * 128 + (ks & 3) * 32) >> 4;
                    umma_f16(tm_dp, dV + ok, dDO + oq, idesc_s, ks ? 1u : 0u);
                }
                umma_commit(s_full);
            }
            __syncwarp();
        };
        mbar_wait(kv_full, 0);
        issue_s_dp(0);
        for (int i = 0; i < n_tiles; ++i) {
            const int st = i % ST;
            MBAR_CHAIN_WAIT(p_full, i & 1);  // P^T / dS^T(i) are in shared memory; S^T / dP^T(i) have been read out
            TL(2048 + i * 4 + 0);
            if (ST == 2 && i + 1 < n_tiles) issue_s_dp(i + 1);
            TL(2048 + i * 4 + 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dQm = desc_mn_sw128(smem_u32(sQ(st)), BQ * 128), dDOm = desc_mn_sw128(smem_u32(sDO(st)), BQ * 128);
                const uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < BQ / 16; ++ks) {  // contraction over the queries of this tile
                    const uint32_t oa = ((ks >> 2) * 128 * 128 + (ks & 3) * 32) >> 4;
                    umma_f16(tm_dv, dPT + oa, dDOm + ((ks * 2048) >> 4), idesc_acc, (acc | ks) ? 1u : 0u);
                    umma_f16(tm_dk, dDST + oa, dQm + ((ks * 2048) >> 4), idesc_acc, (acc | ks) ? 1u : 0u);
                }
                umma_commit(&q_free[st]);
                if (i + 1 == n_tiles) umma_commit(acc_done);
            }
            __syncwarp();
            TL(2048 + i * 4 + 2);
            if (ST == 1 && i + 1 < n_tiles) issue_s_dp(i + 1);
        }
    } else {
        const int r = warp * 32 + lane;  // key row of this thread
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const bool key_ok = k0 + r < p.Nk;
        const long long stat_base = (static_cast<long long>(img) * p.heads + head) * p.Nq;
        // lse / delta of a query tile are fetched one tile ahead (registers) and parked in the other stats buffer at the
        // end of the iteration: the global-load latency used to sit in front of every tile's bar.sync (ncu: the two
        // stats stores + the barrier were 15 % of all stall samples)
        float nl = 0.f, nd = 0.f;
        if (r < BQ) {
            sStat[r] = r < p.Nq ? p.lse[stat_base + r] : 0.f;
            sStat[BQ + r] = r < p.Nq ? p.delta[stat_base + r] : 0.f;
        }
        for (int i = 0; i < n_tiles; ++i) {
            float* sLse = sStat + (i & 1) * 2 * BQ;
            float* sDel = sLse + BQ;
            // stats(i) visible; everyone is done with tile i - 1, so the other buffer may be rewritten at the end of this tile
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (r < BQ && i + 1 < n_tiles) {
                const int q = (i + 1) * BQ + r;
                nl = q < p.Nq ? p.lse[stat_base + q] : 0.f;
                nd = q < p.Nq ? p.delta[stat_base + q] : 0.f;
            }
            if (warp == 0) TL(i * 8 + 0);
            MBAR_CHAIN_WAIT(s_full, i & 1);
            if (warp == 0) TL(i * 8 + 1);
            tc_fence_after();
            const int q_valid = min(BQ, p.Nq - i * BQ);
#pragma unroll 1
            for (int c = 0; c < BQ; c += 32) {
                uint32_t sr[32], dr[32], pp[16], dd[16];
                tmem_ld_32x32(tm_s + lane_off + c, sr);
                tmem_ld_32x32(tm_dp + lane_off + c, dr);
                tmem_ld_wait();
                if (warp == 0) TL(i * 8 + 2 + (c >> 5) * 2);
                // the dV / dK MMAs of the previous tile read P^T / dS^T: done before these are overwritten
                if (c == 0 && i > 0) mbar_wait(&q_free[(i - 1) % ST], ((i - 1) / ST) & 1);
                if (warp == 0) TL(i * 8 + 3 + (c >> 5) * 2);
                float ls[32], de[32];
#pragma unroll
                for (int t = 0; t < 32; t += 4) {  // broadcast 16-byte shared loads: every lane reads the same columns
                    const float4 a4 = lds128f(smem_u32(sLse + c + t));
                    const float4 b4 = lds128f(smem_u32(sDel + c + t));
                    ls[t] = a4.x; ls[t + 1] = a4.y; ls[t + 2] = a4.z; ls[t + 3] = a4.w;
                    de[t] = b4.x; de[t + 1] = b4.y; de[t + 2] = b4.z; de[t + 3] = b4.w;
                }
                if (key_ok && q_valid == BQ) {
#pragma unroll
                    for (int t = 0; t < 32; t += 2) {
                        const float p0 = fast_exp2(fmaf(__uint_as_float(sr[t]), p.scale_log2e, -ls[t]));
                        const float p1 = fast_exp2(fmaf(__uint_as_float(sr[t + 1]), p.scale_log2e, -ls[t + 1]));
                        pp[t >> 1] = pack_half2(p0, p1);
                        dd[t >> 1] = pack_half2(p0 * (__uint_as_float(dr[t]) - de[t]), p1 * (__uint_as_float(dr[t + 1]) - de[t + 1]));
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 32; t += 2) {
                        float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
                        if (key_ok && c + t < q_valid) {
                            p0 = fast_exp2(fmaf(__uint_as_float(sr[t]), p.scale_log2e, -ls[t]));
                            d0 = p0 * (__uint_as_float(dr[t]) - de[t]);
                        }
                        if (key_ok && c + t + 1 < q_valid) {
                            p1 = fast_exp2(fmaf(__uint_as_float(sr[t + 1]), p.scale_log2e, -ls[t + 1]));
                            d1 = p1 * (__uint_as_float(dr[t + 1]) - de[t + 1]);
                        }
                        pp[t >> 1] = pack_half2(p0, p1);
                        dd[t >> 1] = pack_half2(d0, d1);
                    }
                }
                store_row_chunk(smem_u32(sPT), 128, r, c, pp);
                store_row_chunk(smem_u32(sDST), 128, r, c, dd);
            }
            if (warp == 0) TL(i * 8 + 6);
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
            if (warp == 0) TL(i * 8 + 7);
            if (r < BQ && i + 1 < n_tiles) {
                float* nb = sStat + ((i + 1) & 1) * 2 * BQ;
                nb[r] = nl;
                nb[BQ + r] = nd;
            }
        }
        mbar_wait(acc_done, 0);
        tc_fence_after();
        __half* ov = p.dv + (static_cast<long long>(img) * p.Nk + k0 + r) * p.lddv + head * p.d;
        __half* okk = p.dk + (static_cast<long long>(img) * p.Nk + k0 + r) * p.lddk + head * p.d;
#pragma unroll
        for (int c = 0; c < DPAD; c += 16) {
            if (c < p.d) {
                uint32_t rv[16], rk[16];
                tmem_ld_32x16(tm_dv + lane_off + c, rv);
                tmem_ld_32x16(tm_dk + lane_off + c, rk);
                tmem_ld_wait();
                if (key_ok) {
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (c + g * 8 < p.d) {
                            uint4 u, w;
                            uint32_t* uu = reinterpret_cast<uint32_t*>(&u);
                            uint32_t* ww = reinterpret_cast<uint32_t*>(&w);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                uu[e] = pack_half2(__uint_as_float(rv[g * 8 + 2 * e]), __uint_as_float(rv[g * 8 + 2 * e + 1]));
                                ww[e] = pack_half2(__uint_as_float(rk[g * 8 + 2 * e]) * p.scale, __uint_as_float(rk[g * 8 + 2 * e + 1]) * p.scale);
                            }
                            *reinterpret_cast<uint4*>(ov + c + g * 8) = u;
                            *reinterpret_cast<uint4*>(okk + c + g * 8) = w;
                        }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// ================================================================================================ dQ, all-TMEM
// Same recipe as the all-TMEM dK/dV kernel below, for d <= 48 and long key sequences: one 512-column CTA per SM,
// S / dP double-buffered, Q / dO / dS in TMEM, 8 row-math warps (two per TMEM lane group, one 32-key half each).
//   columns: S0 64 | dP0 64 | S1 64 | dP1 64 | dQ 48 | dS 32 | Q 24 | dO 24 = 384.
struct DqTsSmem {
    static constexpr int BKV = 64, ST = 4;
    static constexpr int Q_BYTES = 128 * 128;           // Q and dO: [128 queries][128 B]
    static constexpr int KV_BYTES = BKV * 128;          // K and V tile: [64 keys][128 B]
    static constexpr int STAGE_BYTES = 2 * KV_BYTES;
    static constexpr int OFF_DO = Q_BYTES, OFF_K = 2 * Q_BYTES;
    static constexpr int DATA = OFF_K + ST * STAGE_BYTES;
    static constexpr int TOTAL = DATA + 1024 + 256;
};

constexpr int DQ_TS_THREADS = 320;  // warps 0-7: row math, 8: TMA, 9: MMA

__global__ void __launch_bounds__(DQ_TS_THREADS, 1)
attn_bwd_dq_ts_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ AttnBwdParams p) {
    using L = DqTsSmem;
    constexpr int BKV = L::BKV, ST = L::ST, DPAD = 48, KS = DPAD / 16;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem, *sDO = smem + L::OFF_DO;
    auto sK = [&](int st) { return smem + L::OFF_K + st * L::STAGE_BYTES; };
    auto sV = [&](int st) { return smem + L::OFF_K + st * L::STAGE_BYTES + L::KV_BYTES; };
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::DATA);
    uint64_t *q_full = bars, *kv_full = bars + 1 /*[4]*/, *kv_free = bars + 5 /*[4]*/, *s_full = bars + 9 /*[2]*/, *ds_full = bars + 11,
             *acc_done = bars + 12, *a_ready = bars + 13;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
    if (warp == 8 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    if (warp == 9 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < ST; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_free[i], 1); }
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        mbar_init(ds_full, 256); mbar_init(acc_done, 1); mbar_init(a_ready, 256);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    const uint32_t tm_sdp = tmem_base;                 // buffer b: S at + 128 b, dP at + 128 b + 64
    const uint32_t tm_dq = tmem_base + 256, tm_ds16 = tm_dq + DPAD, tm_q16 = tm_ds16 + BKV / 2, tm_do16 = tm_q16 + DPAD / 2;
    const int n_tiles = (p.Nk + BKV - 1) / BKV;

    if (warp == 8) {
        if (lane == 0) {
            mbar_expect_tx(q_full, 2 * 128 * 128);
            tma_load_4d(sQ, &tmQ, q_full, 0, head, q0, img);
            tma_load_4d(sDO, &tmDO, q_full, 0, head, q0, img);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % ST;
                if (j >= ST) mbar_wait(&kv_free[st], ((j / ST) - 1) & 1);
                mbar_expect_tx(&kv_full[st], 2 * BKV * 128);
                tma_load_4d(sK(st), &tmK, &kv_full[st], 0, head, j * BKV, img);
                tma_load_4d(sV(st), &tmV, &kv_full[st], 0, head, j * BKV, img);
            }
        }
    } else if (warp == 9) {
        const uint32_t idesc_s = p.idesc_s, idesc_acc = p.idesc_acc;
        auto issue_s_dp = [&](int j) {  // S = Q K(j)^T, dP = dO V(j)^T into buffer j & 1
            const int st = j % ST;
            mbar_wait(&kv_full[st], (j / ST) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(sK(st))), dV = umma_desc_kmajor_sw128(smem_u32(sV(st)));
                const uint32_t ts = tm_sdp + 128 * (j & 1);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) umma_f16_ts(ts, tm_q16 + 8 * ks, dK + 2 * ks, idesc_s, ks ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) umma_f16_ts(ts + 64, tm_do16 + 8 * ks, dV + 2 * ks, idesc_s, ks ? 1u : 0u);
                umma_commit(&s_full[j & 1]);
            }
            __syncwarp();
        };
        mbar_wait(a_ready, 0);  // Q / dO rows are in TMEM
        issue_s_dp(0);
        if (n_tiles > 1) issue_s_dp(1);
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j % ST;
            mbar_wait(ds_full, j & 1);  // dS(j) is in TMEM; buffer j & 1 of S / dP has been read out
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dKm = desc_mn_sw128(smem_u32(sK(st)), BKV * 128);
                const uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks)  // contraction over the keys of this tile
                    umma_f16_ts(tm_dq, tm_ds16 + 8 * ks, dKm + ((ks * 2048) >> 4), idesc_acc, (acc | ks) ? 1u : 0u);
                umma_commit(&kv_free[st]);  // K / V stage and the dS columns are free once these complete
                if (j + 1 == n_tiles) umma_commit(acc_done);
            }
            __syncwarp();
            if (j + 2 < n_tiles) issue_s_dp(j + 2);
        }
    } else {
        const int lg = warp & 3, part = warp >> 2;  // TMEM lane group; 32-key half of every tile
        const int r = lg * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(lg * 32) << 16;
        const bool row_ok = q0 + r < p.Nq;
        const long long stat_idx = (static_cast<long long>(img) * p.heads + head) * p.Nq + q0 + r;
        const float lse = row_ok ? p.lse[stat_idx] : 0.f;
        const float dl = row_ok ? p.delta[stat_idx] : 0.f;
        {
            // this query's Q (part 0) / dO (part 1) row -> TMEM (SWIZZLE_128B tile: unit u sits at u ^ (r & 7))
            mbar_wait(q_full, 0);
            const uint32_t src = smem_u32(part == 0 ? sQ : sDO) + r * 128;
            const uint32_t dst = (part == 0 ? tm_q16 : tm_do16) + lane_off;
            uint32_t w[24];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const uint4 x = lds128(src + ((u ^ (r & 7)) << 4));
                w[4 * u] = x.x; w[4 * u + 1] = x.y; w[4 * u + 2] = x.z; w[4 * u + 3] = x.w;
            }
            tmem_st_32x16(dst, w);
            tmem_st_32x8(dst + 16, w + 16);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(a_ready);
        }
        const int c = 32 * part;
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const int kv_valid = min(BKV, p.Nk - j * BKV);
            const uint32_t ts = tm_sdp + 128 * (j & 1) + lane_off;
            uint32_t sr[32], dr[32], packed[16];
            tmem_ld_32x32(ts + c, sr);
            tmem_ld_32x32(ts + 64 + c, dr);
            tmem_ld_wait();
            if (kv_valid == BKV) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float p0 = fast_exp2(fmaf(__uint_as_float(sr[i]), p.scale_log2e, -lse));
                    const float p1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), p.scale_log2e, -lse));
                    packed[i >> 1] = pack_half2(p0 * (__uint_as_float(dr[i]) - dl), p1 * (__uint_as_float(dr[i + 1]) - dl));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float ds0 = 0.f, ds1 = 0.f;
                    if (c + i < kv_valid) ds0 = fast_exp2(fmaf(__uint_as_float(sr[i]), p.scale_log2e, -lse)) * (__uint_as_float(dr[i]) - dl);
                    if (c + i + 1 < kv_valid) ds1 = fast_exp2(fmaf(__uint_as_float(sr[i + 1]), p.scale_log2e, -lse)) * (__uint_as_float(dr[i + 1]) - dl);
                    packed[i >> 1] = pack_half2(ds0, ds1);
                }
            }
            // the dQ MMAs of the previous tile read the dS columns: done before they are overwritten
            if (j > 0) mbar_wait(&kv_free[(j - 1) % ST], ((j - 1) / ST) & 1);
            tmem_st_32x16(tm_ds16 + lane_off + (c >> 1), packed);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(ds_full);
        }
        mbar_wait(acc_done, 0);
        tc_fence_after();
        __half* out = p.dq + (static_cast<long long>(img) * p.Nq + q0 + r) * p.lddq + head * p.d;
#pragma unroll
        for (int cc = 0; cc < DPAD; cc += 16) {
            if (cc < p.d && ((cc >> 4) & 1) == part) {
                uint32_t raw[16];
                tmem_ld_32x16(tm_dq + lane_off + cc, raw);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (cc + g * 8 < p.d) {
                            uint4 u;
                            uint32_t* w = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                w[e] = pack_half2(__uint_as_float(raw[g * 8 + 2 * e]) * p.scale, __uint_as_float(raw[g * 8 + 2 * e + 1]) * p.scale);
                            *reinterpret_cast<uint4*>(out + cc + g * 8) = u;
                        }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

static int launch_dq_ts(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& tk, const CUtensorMap& tv,
                        const AttnBwdParams& p, dim3 grid, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_bwd_dq_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DqTsSmem::TOTAL) != cudaSuccess)
            return CTRLORA_ERR_CUDA;
        attr = true;
    }
    return launch_pdl(attn_bwd_dq_ts_kernel, grid, dim3(DQ_TS_THREADS), (size_t)DqTsSmem::TOTAL, s, tq, tdo, tk, tv, p) == cudaSuccess
               ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

// ================================================================================================ dK, dV, all-TMEM
// d <= 48 (the 64x64 level: almost all of the backward attention time). One CTA per SM with the whole 512-column TMEM:
//   S^T / dP^T double-buffered (2 x 128 columns): the products of query tile i+1 (and i+2) are computed while the row
//     threads work on tile i -- the row math never waits for the tensor pipe;
//   every A operand in TMEM: K and V rows copied once (thread = key row), P^T and dS^T written by the row math
//     (tcgen05.st) -- an A-from-TMEM MMA costs N/2 cycles per k-step instead of 32 + N/4 (tools/microbench/mma_issue.cu):
//     384 instead of 640 tensor cycles per 128 x 64 tile, and no shared-memory round trip for P^T / dS^T.
//   columns: S^T0 64 | dP^T0 64 | S^T1 64 | dP^T1 64 | dV 48 | dK 48 | P^T 32 | dS^T 32 | K 24 | V 24 = 464.
struct DkvTsSmem {
    static constexpr int BQ = 64, ST = 4;
    static constexpr int KV_BYTES = 128 * 128;          // K and V: [128 keys][128 B] (one 64-column chunk, d <= 48)
    static constexpr int Q_BYTES = BQ * 128;            // Q and dO tile: [64 queries][128 B]
    static constexpr int STAGE_BYTES = 2 * Q_BYTES;
    static constexpr int OFF_V = KV_BYTES, OFF_Q = 2 * KV_BYTES, OFF_STAT = OFF_Q + ST * STAGE_BYTES;
    static constexpr int DATA = OFF_STAT + 2 * 2 * BQ * 4;
    static constexpr int BAR = (DATA + 127) / 128 * 128;
    static constexpr int TOTAL = BAR + 1024 + 256;
};

constexpr int DKV_TS_THREADS = 320;  // warps 0-7: row math (lane group = warp & 3, 32-query half = warp >> 2), 8: TMA, 9: MMA

__global__ void __launch_bounds__(DKV_TS_THREADS, 1)
attn_bwd_dkdv_ts_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                        const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                        const __grid_constant__ AttnBwdParams p) {
    using L = DkvTsSmem;
    constexpr int BQ = L::BQ, ST = L::ST, DPAD = 48, KS = DPAD / 16;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sK = smem, *sV = smem + L::OFF_V;
    auto sQ = [&](int st) { return smem + L::OFF_Q + st * L::STAGE_BYTES; };
    auto sDO = [&](int st) { return smem + L::OFF_Q + st * L::STAGE_BYTES + L::Q_BYTES; };
    float* sStat = reinterpret_cast<float*>(smem + L::OFF_STAT);  // [2 buffers][lse BQ | delta BQ]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR);
    uint64_t *kv_full = bars, *q_full = bars + 1 /*[4]*/, *q_free = bars + 5 /*[4]*/, *s_full = bars + 9 /*[2]*/, *p_full = bars + 11,
             *acc_done = bars + 12, *a_ready = bars + 13;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int k0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
    if (warp == 8 && lane == 0) { tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); }
    if (warp == 9 && lane == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < ST; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_free[i], 1); }
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
        mbar_init(p_full, 256); mbar_init(acc_done, 1); mbar_init(a_ready, 256);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    const uint32_t tm_sdp = tmem_base;                 // buffer b: S^T at + 128 b, dP^T at + 128 b + 64
    const uint32_t tm_dv = tmem_base + 256, tm_dk = tm_dv + DPAD;
    const uint32_t tm_pt = tm_dk + DPAD, tm_dst = tm_pt + BQ / 2, tm_k16 = tm_dst + BQ / 2, tm_v16 = tm_k16 + DPAD / 2;
    const int n_tiles = (p.Nq + BQ - 1) / BQ;

    if (warp == 8) {
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * 128 * 128);
            tma_load_4d(sK, &tmK, kv_full, 0, head, k0, img);
            tma_load_4d(sV, &tmV, kv_full, 0, head, k0, img);
            for (int i = 0; i < n_tiles; ++i) {
                const int st = i % ST;
                if (i >= ST) mbar_wait(&q_free[st], ((i / ST) - 1) & 1);
                mbar_expect_tx(&q_full[st], 2 * BQ * 128);
                tma_load_4d(sQ(st), &tmQ, &q_full[st], 0, head, i * BQ, img);
                tma_load_4d(sDO(st), &tmDO, &q_full[st], 0, head, i * BQ, img);
            }
        }
    } else if (warp == 9) {
        const uint32_t idesc_s = p.idesc_s, idesc_acc = p.idesc_acc;
        auto issue_s_dp = [&](int i) {  // S^T = K Q(i)^T, dP^T = V dO(i)^T into buffer i & 1
            const int st = i % ST;
            mbar_wait(&q_full[st], (i / ST) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dQ = umma_desc_kmajor_sw128(smem_u32(sQ(st))), dDO = umma_desc_kmajor_sw128(smem_u32(sDO(st)));
                const uint32_t ts = tm_sdp + 128 * (i & 1);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) umma_f16_ts(ts, tm_k16 + 8 * ks, dQ + 2 * ks, idesc_s, ks ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) umma_f16_ts(ts + 64, tm_v16 + 8 * ks, dDO + 2 * ks, idesc_s, ks ? 1u : 0u);
                umma_commit(&s_full[i & 1]);
            }
            __syncwarp();
        };
        mbar_wait(a_ready, 0);  // K / V rows are in TMEM
        issue_s_dp(0);
        if (n_tiles > 1) issue_s_dp(1);
        for (int i = 0; i < n_tiles; ++i) {
            const int st = i % ST;
            mbar_wait(p_full, i & 1);  // P^T / dS^T(i) are in TMEM; buffer i & 1 of S^T / dP^T has been read out
            TL(2048 + i * 4 + 0);
            tc_fence_after();
            if (elect_one()) {
                const uint64_t dQm = desc_mn_sw128(smem_u32(sQ(st)), BQ * 128), dDOm = desc_mn_sw128(smem_u32(sDO(st)), BQ * 128);
                const uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < BQ / 16; ++ks) {  // contraction over the queries of this tile
                    umma_f16_ts(tm_dv, tm_pt + 8 * ks, dDOm + ((ks * 2048) >> 4), idesc_acc, (acc | ks) ? 1u : 0u);
                    umma_f16_ts(tm_dk, tm_dst + 8 * ks, dQm + ((ks * 2048) >> 4), idesc_acc, (acc | ks) ? 1u : 0u);
                }
                umma_commit(&q_free[st]);  // Q / dO stage and the P^T / dS^T columns are free once these complete
                if (i + 1 == n_tiles) umma_commit(acc_done);
            }
            __syncwarp();
            TL(2048 + i * 4 + 1);
            if (i + 2 < n_tiles) issue_s_dp(i + 2);
            TL(2048 + i * 4 + 2);
        }
    } else {
        // two warps per TMEM lane group: each takes one 32-query half of every tile (nothing couples the columns in
        // the backward), so every scheduler has two row-math warps to overlap TMEM / MUFU latency
        const int lg = warp & 3, part = warp >> 2;
        const int r = lg * 32 + lane;  // key row of this thread
        const uint32_t lane_off = static_cast<uint32_t>(lg * 32) << 16;
        const bool key_ok = k0 + r < p.Nk;
        const long long stat_base = (static_cast<long long>(img) * p.heads + head) * p.Nq;
        {
            // this key's K and V rows -> TMEM (SWIZZLE_128B tile: 16-byte unit u sits at u ^ (r & 7); columns >= d are zero)
            mbar_wait(kv_full, 0);
            const uint32_t src = smem_u32(part == 0 ? sK : sV) + r * 128;  // part 0 copies K, part 1 copies V
            const uint32_t dst = (part == 0 ? tm_k16 : tm_v16) + lane_off;
            uint32_t w[24];
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const uint4 x = lds128(src + ((u ^ (r & 7)) << 4));
                w[4 * u] = x.x; w[4 * u + 1] = x.y; w[4 * u + 2] = x.z; w[4 * u + 3] = x.w;
            }
            tmem_st_32x16(dst, w);
            tmem_st_32x8(dst + 16, w + 16);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(a_ready);
        }
        float nl = 0.f, nd = 0.f;
        if (part == 0 && r < BQ) {
            sStat[r] = r < p.Nq ? p.lse[stat_base + r] : 0.f;
            sStat[BQ + r] = r < p.Nq ? p.delta[stat_base + r] : 0.f;
        }
        for (int i = 0; i < n_tiles; ++i) {
            float* sLse = sStat + (i & 1) * 2 * BQ;
            float* sDel = sLse + BQ;
            asm volatile("bar.sync 1, 256;" ::: "memory");  // stats(i) visible; the other buffer is free
            if (part == 0 && r < BQ && i + 1 < n_tiles) {
                const int q = (i + 1) * BQ + r;
                nl = q < p.Nq ? p.lse[stat_base + q] : 0.f;
                nd = q < p.Nq ? p.delta[stat_base + q] : 0.f;
            }
            if (warp == 0) TL(i * 8 + 0);
            mbar_wait(&s_full[i & 1], (i >> 1) & 1);
            if (warp == 0) TL(i * 8 + 1);
            tc_fence_after();
            const int q_valid = min(BQ, p.Nq - i * BQ);
            const uint32_t ts = tm_sdp + 128 * (i & 1) + lane_off;
            const bool full = key_ok && q_valid == BQ;
            {
                const int c = 32 * part;
                uint32_t sr[32], dr[32], pp[16], dd[16];
                tmem_ld_32x32(ts + c, sr);
                tmem_ld_32x32(ts + 64 + c, dr);
                tmem_ld_wait();
                if (warp == 0) TL(i * 8 + 2 + (c >> 5) * 2);
                float ls[32], de[32];
#pragma unroll
                for (int t = 0; t < 32; t += 4) {  // broadcast 16-byte shared loads: every lane reads the same columns
                    const float4 a4 = lds128f(smem_u32(sLse + c + t));
                    const float4 b4 = lds128f(smem_u32(sDel + c + t));
                    ls[t] = a4.x; ls[t + 1] = a4.y; ls[t + 2] = a4.z; ls[t + 3] = a4.w;
                    de[t] = b4.x; de[t + 1] = b4.y; de[t + 2] = b4.z; de[t + 3] = b4.w;
                }
                if (full) {  // straight-line: the masked form below compiles to a branch per element
#pragma unroll
                    for (int t = 0; t < 32; t += 2) {
                        const float p0 = fast_exp2(fmaf(__uint_as_float(sr[t]), p.scale_log2e, -ls[t]));
                        const float p1 = fast_exp2(fmaf(__uint_as_float(sr[t + 1]), p.scale_log2e, -ls[t + 1]));
                        pp[t >> 1] = pack_half2(p0, p1);
                        dd[t >> 1] = pack_half2(p0 * (__uint_as_float(dr[t]) - de[t]), p1 * (__uint_as_float(dr[t + 1]) - de[t + 1]));
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 32; t += 2) {
                        float p0 = 0.f, p1 = 0.f, d0 = 0.f, d1 = 0.f;
                        if (key_ok && c + t < q_valid) {
                            p0 = fast_exp2(fmaf(__uint_as_float(sr[t]), p.scale_log2e, -ls[t]));
                            d0 = p0 * (__uint_as_float(dr[t]) - de[t]);
                        }
                        if (key_ok && c + t + 1 < q_valid) {
                            p1 = fast_exp2(fmaf(__uint_as_float(sr[t + 1]), p.scale_log2e, -ls[t + 1]));
                            d1 = p1 * (__uint_as_float(dr[t + 1]) - de[t + 1]);
                        }
                        pp[t >> 1] = pack_half2(p0, p1);
                        dd[t >> 1] = pack_half2(d0, d1);
                    }
                }
                // the dV / dK MMAs of the previous tile read the P^T / dS^T columns: done before these are overwritten
                if (warp == 0) TL(i * 8 + 3 + (c >> 5) * 2);
                if (i > 0) mbar_wait(&q_free[(i - 1) % ST], ((i - 1) / ST) & 1);
                tmem_st_32x16(tm_pt + lane_off + (c >> 1), pp);
                tmem_st_32x16(tm_dst + lane_off + (c >> 1), dd);
            }
            if (warp == 0) TL(i * 8 + 6);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_full);
            if (warp == 0) TL(i * 8 + 7);
            if (part == 0 && r < BQ && i + 1 < n_tiles) {
                float* nb = sStat + ((i + 1) & 1) * 2 * BQ;
                nb[r] = nl;
                nb[BQ + r] = nd;
            }
        }
        mbar_wait(acc_done, 0);
        tc_fence_after();
        __half* ov = p.dv + (static_cast<long long>(img) * p.Nk + k0 + r) * p.lddv + head * p.d;
        __half* okk = p.dk + (static_cast<long long>(img) * p.Nk + k0 + r) * p.lddk + head * p.d;
        const float osc = part == 0 ? 1.0f : p.scale;  // part 0 stores dV, part 1 stores dK (scaled by d^-1/2)
        __half* orow = part == 0 ? ov : okk;
        const uint32_t tsrc = (part == 0 ? tm_dv : tm_dk) + lane_off;
#pragma unroll
        for (int c = 0; c < DPAD; c += 16) {
            if (c < p.d) {
                uint32_t rv[16];
                tmem_ld_32x16(tsrc + c, rv);
                tmem_ld_wait();
                if (key_ok) {
#pragma unroll
                    for (int g = 0; g < 2; ++g)
                        if (c + g * 8 < p.d) {
                            uint4 u;
                            uint32_t* uu = reinterpret_cast<uint32_t*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                uu[e] = pack_half2(__uint_as_float(rv[g * 8 + 2 * e]) * osc, __uint_as_float(rv[g * 8 + 2 * e + 1]) * osc);
                            *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
                        }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc(tmem_base, 512); }
}

static int launch_dkdv_ts(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& tk, const CUtensorMap& tv,
                          const AttnBwdParams& p, dim3 grid, cudaStream_t s) {
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_bwd_dkdv_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvTsSmem::TOTAL) != cudaSuccess)
            return CTRLORA_ERR_CUDA;
        attr = true;
    }
    return launch_pdl(attn_bwd_dkdv_ts_kernel, grid, dim3(DKV_TS_THREADS), (size_t)DkvTsSmem::TOTAL, s, tq, tdo, tk, tv, p) == cudaSuccess
               ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

template <int DPAD, int BKV>
static int launch_dq(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& tk, const CUtensorMap& tv,
                     const AttnBwdParams& p, dim3 grid, cudaStream_t s) {
    using L = DqSmem<DPAD, BKV>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_bwd_dq_kernel<DPAD, BKV>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess)
            return CTRLORA_ERR_CUDA;
        attr = true;
    }
    return launch_pdl(attn_bwd_dq_kernel<DPAD, BKV>, grid, dim3(AB_THREADS), (size_t)L::TOTAL, s, tq, tdo, tk, tv, p) == cudaSuccess
               ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

template <int DPAD, int BQ>
static int launch_dkdv(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& tk, const CUtensorMap& tv,
                       const AttnBwdParams& p, dim3 grid, cudaStream_t s) {
    using L = DkvSmem<DPAD, BQ>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attn_bwd_dkdv_kernel<DPAD, BQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess)
            return CTRLORA_ERR_CUDA;
        attr = true;
    }
    return launch_pdl(attn_bwd_dkdv_kernel<DPAD, BQ>, grid, dim3(AB_THREADS), (size_t)L::TOTAL, s, tq, tdo, tk, tv, p) == cudaSuccess
               ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

static int tmap_tokens(CUtensorMap* m, const void* base, long long ld, int d, int heads, int n, int batch, int box_rows) {
    uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)n, (uint64_t)batch};
    uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ld * 2, (uint64_t)ld * 2 * n};
    uint32_t box[4] = {64, 1, (uint32_t)box_rows, 1};
    return make_tmap_f16(m, base, 4, dims, str, box);
}

}  // namespace ctrl

using namespace ctrl;

extern "C" int ctrlora_attention_bwd_f16(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                         long long ldv, const void* o, long long ldo, const void* dout, long long lddo,
                                         const float* lse, float* delta_ws, void* dq, long long lddq, void* dk, long long lddk,
                                         void* dv, long long lddv, int batch, int heads, int nq, int nk, int head_dim,
                                         void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!q || !k || !v || !o || !dout || !lse || !delta_ws || !dq || !dk || !dv) return CTRLORA_ERR_ARG;
    const int d = head_dim;
    if (d % 8 || d > 160 || ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || lddo % 8 || lddq % 8 || lddk % 8 || lddv % 8)
        return CTRLORA_ERR_ARG;
    {
        const long long total = static_cast<long long>(batch) * nq * heads;
        launch_pdl(attn_bwd_delta_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)0, stream,
                   reinterpret_cast<const __half*>(o), ldo, reinterpret_cast<const __half*>(dout), lddo, delta_ws, batch, heads, nq, d);
    }
    AttnBwdParams p;
    memset(&p, 0, sizeof(p));
    p.Nq = nq; p.Nk = nk; p.heads = heads; p.d = d; p.d16 = (d + 15) / 16 * 16; p.nkc = (d + 63) / 64;
    p.scale = 1.0f / sqrtf(static_cast<float>(d));
    p.scale_log2e = p.scale * 1.4426950408889634f;
    p.lse = lse; p.delta = delta_ws;
    p.dq = reinterpret_cast<__half*>(dq); p.lddq = lddq;
    p.dk = reinterpret_cast<__half*>(dk); p.lddk = lddk;
    p.dv = reinterpret_cast<__half*>(dv); p.lddv = lddv;
    p.idesc_acc = umma_idesc_f16(128, p.d16, 0) | (1u << 16);  // B operand MN-major
    const bool wide = d > 80;
    // ---- dQ: 128-query CTAs, key tiles of 128 (64 for d_head 160)
    {
        const int bkv = (wide || d <= 48) ? 64 : 128;  // d <= 48: 64-key tiles keep TMEM at 256 columns -> 2 CTAs per SM
        CUtensorMap tq, tdo, tk, tv;
        int rc = tmap_tokens(&tq, q, ldq, d, heads, nq, batch, 128);
        if (!rc) rc = tmap_tokens(&tdo, dout, lddo, d, heads, nq, batch, 128);
        if (!rc) rc = tmap_tokens(&tk, k, ldk, d, heads, nk, batch, bkv);
        if (!rc) rc = tmap_tokens(&tv, v, ldv, d, heads, nk, batch, bkv);
        if (rc) return rc;
        p.idesc_s = umma_idesc_f16(128, bkv, 0);
        dim3 grid((nq + 127) / 128, heads, batch);
        static int tsq_env = -1;
        if (tsq_env < 0) {
            const char* e = getenv("CTRLORA_ATTN_BWD_TS");
            tsq_env = (e && e[0] == '0') ? 0 : 1;
        }
        if (d <= 48 && tsq_env && nk >= 512) rc = launch_dq_ts(tq, tdo, tk, tv, p, grid, stream);
        else if (d <= 48) rc = launch_dq<48, 64>(tq, tdo, tk, tv, p, grid, stream);
        else if (d <= 80) rc = launch_dq<80, 128>(tq, tdo, tk, tv, p, grid, stream);
        else rc = launch_dq<160, 64>(tq, tdo, tk, tv, p, grid, stream);
        if (rc) return rc;
    }
    // ---- dK, dV: 128-key CTAs, query tiles of 128 (64 for d_head 160)
    {
        const int bq = (wide || d <= 48) ? 64 : 128;
        CUtensorMap tq, tdo, tk, tv;
        int rc = tmap_tokens(&tq, q, ldq, d, heads, nq, batch, bq);
        if (!rc) rc = tmap_tokens(&tdo, dout, lddo, d, heads, nq, batch, bq);
        if (!rc) rc = tmap_tokens(&tk, k, ldk, d, heads, nk, batch, 128);
        if (!rc) rc = tmap_tokens(&tv, v, ldv, d, heads, nk, batch, 128);
        if (rc) return rc;
        p.idesc_s = umma_idesc_f16(128, bq, 0);
        dim3 grid((nk + 127) / 128, heads, batch);
        static int ts_env = -1;
        if (ts_env < 0) {
            const char* e = getenv("CTRLORA_ATTN_BWD_TS");
            ts_env = (e && e[0] == '0') ? 0 : 1;  // 0: the two-CTAs-per-SM shared-memory-operand kernel (kept for A/B runs)
        }
        if (d <= 48 && ts_env && nk >= 512) rc = launch_dkdv_ts(tq, tdo, tk, tv, p, grid, stream);
        else if (d <= 48) rc = launch_dkdv<48, 64>(tq, tdo, tk, tv, p, grid, stream);
        else if (d <= 80) rc = launch_dkdv<80, 128>(tq, tdo, tk, tv, p, grid, stream);
        else rc = launch_dkdv<160, 64>(tq, tdo, tk, tv, p, grid, stream);
        if (rc) return rc;
    }
    return cudaGetLastError() == cudaSuccess ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

#ifdef CTRLORA_TIMELINE
extern "C" int ctrlora_debug_timeline(long long* host_out, int n) {
    return cudaMemcpyFromSymbol(host_out, ctrl::g_tl, sizeof(long long) * n) == cudaSuccess ? 0 : 1;
}
#endif
