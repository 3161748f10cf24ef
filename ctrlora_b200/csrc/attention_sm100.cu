// Fused attention forward on tcgen05 for sm_100a:  O = softmax(Q K^T * d^-1/2) V  per (image, head, 128-query tile).
// reference: CrossAttention.forward, ldm/modules/attention.py:163-194 (fp32 logits and softmax, scale d_head^-0.5);
// the [8B, N, N] fp32 `sim` matrix the reference materialises (512 MiB per image at N = 4096) never exists here.
//
//   S = Q K^T      tcgen05.mma, operands staged by TMA (Q [128 x d], K [BKV x d], K-major, SWIZZLE_128B), S in TMEM
//   softmax        128 threads, one query row each (TMEM lane == thread): no cross-thread reductions at all;
//                  fp32 max / exp2 / sum, P written to shared memory as fp16 in the UMMA K-major swizzled layout
//   PV = P V       tcgen05.mma with A = P (smem), B = V^T tile ([d x BKV], K-major; the V projection GEMM stores V
//                  transposed for exactly this), result in TMEM
// MULTI (long key sequences, d <= 80): online softmax, running output kept in registers, rescaled per KV tile.
// single-tile (Nk <= BKV, any d <= 160): exact two-pass softmax, P aliases the dead K buffer.
#include "common.cuh"
#include "ctrlora_b200.h"
#include "gemm_sm100.cuh"
#include <stdlib.h>
#include <string.h>

namespace ctrl {

int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

struct AttnParams {
    int Nq, Nk, heads, d, d16;  // d16 = d rounded up to 16 (UMMA N of the PV product)
    int nkc;                    // K-chunks of 64 over d
    int n_kv_tiles;
    float scale_log2e;          // d^-1/2 * log2(e)
    __half* out;
    long long ldo;
    float* lse;                 // optional [B, H, Nq]: log2-domain log-sum-exp, for the backward
    uint32_t idesc_s, idesc_pv, idesc_l;
};

constexpr int ATT_THREADS = 192;  // warps 0-3: softmax / epilogue (TMEM lane groups 0-3), warp 4: TMA, warp 5: MMA

template <int DPAD, int BKV, bool MULTI>
struct AttnSmem {
    static constexpr int NKC = (DPAD + 63) / 64;
    static constexpr int Q_BYTES = NKC * 128 * 128;
    static constexpr int K_BYTES = NKC * BKV * 128;
    static constexpr int V_CHUNK = DPAD * 128;
    static constexpr int V_BYTES = (BKV / 64) * V_CHUNK;
    static constexpr int P_BYTES = (BKV / 64) * 128 * 128;
    static constexpr int STAGES = MULTI ? 2 : 1;
    static constexpr int STAGE_BYTES = K_BYTES + V_BYTES;
    static constexpr int P_OFF = MULTI ? (Q_BYTES + STAGES * STAGE_BYTES) : Q_BYTES;  // single tile: P aliases K
    static constexpr int DATA_BYTES = MULTI ? (P_OFF + P_BYTES)
                                            : (Q_BYTES + (K_BYTES > P_BYTES ? K_BYTES : P_BYTES) + V_BYTES);
    static constexpr int V_OFF_SINGLE = Q_BYTES + (K_BYTES > P_BYTES ? K_BYTES : P_BYTES);
    // [16 rows x 128 B] of fp16 1.0 (row sums on the tensor core).  Single-tile mode: aliases the Q buffer, which is
    // dead once S = Q K^T has completed (written by the softmax warps after s_full).
    static constexpr int ONES_OFF = MULTI ? DATA_BYTES : 0;
    static constexpr int BAR_OFF = MULTI ? DATA_BYTES + 2048 : DATA_BYTES;
    static constexpr int TOTAL = BAR_OFF + 1024 + 128;
    static constexpr int TMEM_COLS = (BKV + DPAD + 16) <= 256 ? 256 : 512;
};

template <int DPAD, int BKV, bool MULTI>
__global__ void __launch_bounds__(ATT_THREADS, (MULTI && DPAD <= 48) ? 2 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
    using L = AttnSmem<DPAD, BKV, MULTI>;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint8_t* sOnes = smem + L::ONES_OFF;
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;   // [2]
    uint64_t* kv_empty = bars + 3;  // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* pv_full = bars + 7;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;

    uint8_t* sQ = smem;
    auto sK = [&](int stage) { return smem + L::Q_BYTES + (MULTI ? stage * L::STAGE_BYTES : 0); };
    auto sV = [&](int stage) { return MULTI ? smem + L::Q_BYTES + stage * L::STAGE_BYTES + L::K_BYTES : smem + L::V_OFF_SINGLE; };
    uint8_t* sP = smem + L::P_OFF;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 5 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_full, 128);
        mbar_init(pv_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, L::TMEM_COLS);
    if (MULTI && threadIdx.x < 128)  // 128 x 16 B = 2 KiB of 1.0h; all elements equal, so the swizzle is irrelevant
        reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    const uint32_t tmem_s = tmem_base;          // [0, BKV)
    const uint32_t tmem_pv = tmem_base + BKV;   // [BKV, BKV + d16)
    const uint32_t tmem_l = tmem_pv + DPAD;     // [BKV + DPAD, +16): row sums of P (P x ones)
    const int n_tiles = p.n_kv_tiles;

    if (warp == 4) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            mbar_expect_tx(q_full, p.nkc * 128 * 128);
            for (int kc = 0; kc < p.nkc; ++kc) tma_load_4d(sQ + kc * 128 * 128, &tmQ, q_full, kc * 64, head, q0, img);
            const uint32_t tx = p.nkc * BKV * 128 + (BKV / 64) * p.d16 * 128;
            for (int j = 0; j < n_tiles; ++j) {
                const int stage = MULTI ? (j & 1) : 0;
                if (MULTI && j >= 2) mbar_wait(&kv_empty[stage], ((j >> 1) - 1) & 1);
                mbar_expect_tx(&kv_full[stage], tx);
                for (int kc = 0; kc < p.nkc; ++kc)
                    tma_load_4d(sK(stage) + kc * BKV * 128, &tmK, &kv_full[stage], kc * 64, head, j * BKV, img);
                for (int c = 0; c < BKV / 64; ++c)
                    tma_load_4d(sV(stage) + c * L::V_CHUNK, &tmV, &kv_full[stage], j * BKV + c * 64, 0, head, img);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ------------------------------------------------ MMA issuer
            const int ksteps_s = (p.d + 15) / 16;
            auto issue_s = [&](int stage) {
                const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK(stage));
                for (int ks = 0; ks < ksteps_s; ++ks) {
                    const uint32_t off_q = (ks >> 2) * 128 * 128 + (ks & 3) * 32;
                    const uint32_t off_k = (ks >> 2) * BKV * 128 + (ks & 3) * 32;
                    umma_f16(tmem_s, umma_desc_kmajor_sw128(qa + off_q), umma_desc_kmajor_sw128(ka + off_k), p.idesc_s,
                             ks != 0 ? 1u : 0u);
                }
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_s(0);
            for (int j = 0; j < n_tiles; ++j) {
                const int stage = MULTI ? (j & 1) : 0;
                mbar_wait(p_full, j & 1);  // P[j] in smem, S[j] read out, PV[j-1] read out
                tc_fence_after();
                const uint32_t pa = smem_u32(sP), va = smem_u32(sV(stage));
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    const uint32_t off_p = (ks >> 2) * 128 * 128 + (ks & 3) * 32;
                    const uint32_t off_v = (ks >> 2) * L::V_CHUNK + (ks & 3) * 32;
                    umma_f16(tmem_pv, umma_desc_kmajor_sw128(pa + off_p), umma_desc_kmajor_sw128(va + off_v), p.idesc_pv,
                             ks != 0 ? 1u : 0u);
                }
                const uint32_t oa = smem_u32(sOnes);
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    const uint32_t off_p = (ks >> 2) * 128 * 128 + (ks & 3) * 32;
                    umma_f16(tmem_l, umma_desc_kmajor_sw128(pa + off_p), umma_desc_kmajor_sw128(oa), p.idesc_l,
                             ks != 0 ? 1u : 0u);
                }
                umma_commit(pv_full);
                if (MULTI) umma_commit(&kv_empty[stage]);
                if (j + 1 < n_tiles) {
                    const int ns = (j + 1) & 1;
                    mbar_wait(&kv_full[ns], ((j + 1) >> 1) & 1);
                    tc_fence_after();
                    issue_s(ns);
                }
            }
        }
    } else {
        // ---------------------------------------------------- softmax + epilogue: thread == query row
        const int r = warp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        float o_reg[MULTI ? DPAD : 1];
        if (MULTI) {
#pragma unroll
            for (int i = 0; i < DPAD; ++i) o_reg[i] = 0.f;
        }
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            if (!MULTI)  // Q is dead now: its first 2 KiB become the ones tile (made visible by the fence before p_full)
                reinterpret_cast<uint4*>(sOnes)[r] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
            const int kv_valid = min(BKV, p.Nk - j * BKV);  // columns >= kv_valid are TMA zero fill: masked out
            const bool full_tile = kv_valid == BKV;         // warp-uniform
            // pass 1: row max (two TMEM loads in flight, 3-input max)
            float m_tile = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < BKV; c += 64) {
                uint32_t ra[32], rb[32];
                tmem_ld_32x32(tmem_s + lane_off + c, ra);
                tmem_ld_32x32(tmem_s + lane_off + c + 32, rb);
                tmem_ld_wait();
                if (full_tile) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) m_tile = max3f(m_tile, __uint_as_float(ra[i]), __uint_as_float(rb[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (c + i < kv_valid) m_tile = fmaxf(m_tile, __uint_as_float(ra[i]));
                        if (c + 32 + i < kv_valid) m_tile = fmaxf(m_tile, __uint_as_float(rb[i]));
                    }
                }
            }
            const float m_new = fmaxf(m_run, m_tile);
            const float alpha = fast_exp2((m_run - m_new) * p.scale_log2e);  // first tile: exp2(-inf) = 0
            m_run = m_new;
            const float neg_ms = -m_new * p.scale_log2e;
            if (MULTI && j > 0) {
                // fold the previous tile's P*V and P*1 (both relative to the previous max), then rescale to the new max
                mbar_wait(pv_full, (j - 1) & 1);
                tc_fence_after();
                uint32_t lraw;
                tmem_ld_32x1(tmem_l + lane_off, lraw);
#pragma unroll
                for (int c = 0; c < DPAD; c += 16) {
                    uint32_t raw[16];
                    tmem_ld_32x16(tmem_pv + lane_off + c, raw);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o_reg[c + i] = (o_reg[c + i] + __uint_as_float(raw[i])) * alpha;
                }
                l_run = (l_run + __uint_as_float(lraw)) * alpha;
            }
            // pass 2: P = exp2(s * scale*log2e - m * scale*log2e) -> fp16, swizzled K-major rows of 128 B.
            // The row sum is NOT accumulated here: it comes out of the tensor core as P x ones (tmem_l).
#pragma unroll 1
            for (int c = 0; c < BKV; c += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(tmem_s + lane_off + c, raw);
                tmem_ld_wait();
                uint32_t packed[16];
                if (full_tile) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float p0 = fast_exp2(fmaf(__uint_as_float(raw[i]), p.scale_log2e, neg_ms));
                        const float p1 = fast_exp2(fmaf(__uint_as_float(raw[i + 1]), p.scale_log2e, neg_ms));
                        packed[i >> 1] = pack_half2(p0, p1);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float p0 = (c + i < kv_valid) ? fast_exp2(fmaf(__uint_as_float(raw[i]), p.scale_log2e, neg_ms)) : 0.f;
                        const float p1 = (c + i + 1 < kv_valid) ? fast_exp2(fmaf(__uint_as_float(raw[i + 1]), p.scale_log2e, neg_ms)) : 0.f;
                        packed[i >> 1] = pack_half2(p0, p1);
                    }
                }
                const uint32_t chunk = smem_u32(sP) + (c >> 6) * 128 * 128 + r * 128;
                const int u0 = (c & 63) >> 3;  // first 16-byte unit of this 32-column group within the 128 B row
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    sts128(chunk + (((u0 + u) ^ (r & 7)) << 4), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // ---- final: fold the last PV, normalise, store
        mbar_wait(pv_full, (n_tiles - 1) & 1);
        tc_fence_after();
        uint32_t lfin;
        tmem_ld_32x1(tmem_l + lane_off, lfin);
        tmem_ld_wait();
        const float l_tot = l_run + __uint_as_float(lfin);
        const float inv_l = 1.0f / l_tot;
        const bool row_ok = (q0 + r) < p.Nq;
        if (p.lse && row_ok)
            p.lse[(static_cast<long long>(img) * p.heads + head) * p.Nq + q0 + r] = m_run * p.scale_log2e + log2f(l_tot);
        __half* orow = p.out + (static_cast<long long>(img) * p.Nq + q0 + r) * p.ldo + head * p.d;
#pragma unroll
        for (int c = 0; c < DPAD; c += 16) {
            if (c < p.d) {  // warp-uniform
                uint32_t raw[16];
                tmem_ld_32x16(tmem_pv + lane_off + c, raw);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float acc = __uint_as_float(raw[i]);
                    if (MULTI) acc += o_reg[c + i];
                    v[i] = acc * inv_l;
                }
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        if (c + g * 8 < p.d) {
                            uint4 u;
                            __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[g * 8 + 2 * e], v[g * 8 + 2 * e + 1]);
                            *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

// =====================================================================================================================
// Streaming kernel for long key sequences (Nk > 256, d <= 80): the 64x64 / 32x32 self-attentions, i.e. almost all of the
// attention time.  Differences from the MULTI path above (measured 35 % of the MUFU bound, profiles/README.md):
//   * O (and the row sums l) stay in TMEM across KV tiles: the P V product accumulates; nothing is folded through
//     registers per tile.  The softmax reference m_ref is LAZY: a tile only moves it (and rescales O, l in TMEM) when
//     its row maximum exceeds m_ref by more than 2^TAU -- after the first tiles that is rare.  P <= 2^TAU stays far
//     inside fp16 range and the result is algebraically the same softmax.
//   * one pass over S per tile (TMEM read once, exp2 straight away against m_ref); a warp that finds a violating row
//     redoes its 32 rows (slow path).
//   * when d is not a multiple of 16 the zero padding row d of the V^T tile is overwritten with 1.0, so column d of
//     O IS the row sum (no separate P x ones product).
//   * the MMA warp issues S(j+1) BEFORE P V(j): the next tile's logits are ready while P V(j) still runs.
//   * d <= 48: P never goes through shared memory. The row threads write it (fp16 pairs, tcgen05.st) into 64 spare TMEM
//     columns and P V is issued in the A-from-TMEM form, which costs N/2 = 24 cycles per k-step instead of the
//     32 + N/4 = 44 of the smem-A form (tools/microbench/mma_issue.cu: the smem A read is the floor for small N).
constexpr float ATT_TAU = 8.0f;
// Every 2*k-th exponential of the fast path can run on the FMA pipe (exp2_poly3). Measured at 4096 x 4096, d = 40: 0 -> 444 us,
// 1/8 -> 443 us, 1/4 -> 452 us, 1/2 -> 522 us: the kernel is issue/latency-bound before it is MUFU-bound, so it is off.
#ifndef CTRLORA_ATT_POLY_EVERY
#define CTRLORA_ATT_POLY_EVERY 0
#endif
constexpr int ATT_POLY_EVERY = CTRLORA_ATT_POLY_EVERY;

template <int DPAD>
struct StreamSmem {
    static constexpr int BKV = 128;
    static constexpr int NKC = (DPAD + 63) / 64;
    static constexpr int Q_BYTES = NKC * 128 * 128;
    static constexpr int K_BYTES = NKC * BKV * 128;
    static constexpr int V_CHUNK = DPAD * 128;
    static constexpr int V_BYTES = (BKV / 64) * V_CHUNK;
    static constexpr int STAGE_BYTES = K_BYTES + V_BYTES;
    static constexpr int P_OFF = Q_BYTES + 2 * STAGE_BYTES;
    static constexpr int P_BYTES = (BKV / 64) * 128 * 128;
    static constexpr int ONES_OFF = P_OFF + P_BYTES;
    static constexpr int BAR_OFF = ONES_OFF + 2048;
    static constexpr int TOTAL = BAR_OFF + 1024 + 128;
    static constexpr int TMEM_COLS = 256;  // S 128 | O DPAD | l 16
};

// P = exp2(s * sl2 + neg) for one 32-column chunk of one row; tmax tracks the raw row maximum over valid columns.
// POLY = k > 0: the odd element of every k-th pair is evaluated on the FMA pipe (exp2_poly4): k = 1 half, k = 2 a quarter
template <bool MASK, int POLY = ATT_POLY_EVERY>
__device__ __forceinline__ void softmax_chunk(const uint32_t (&raw)[32], float sl2, float neg, int c, int kv_valid,
                                              uint32_t (&packed)[16], float& tmax) {
    if (!MASK) {
        float tm2 = -INFINITY;  // two independent maximum chains
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            if (i & 2) tm2 = max3f(tm2, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
            else tmax = max3f(tmax, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
            const float p0 = fast_exp2(fmaf(__uint_as_float(raw[i]), sl2, neg));
            const float x1 = fmaf(__uint_as_float(raw[i + 1]), sl2, neg);
            const float p1 = (POLY > 0 && ((i >> 1) % (POLY > 0 ? POLY : 1)) == POLY - 1) ? exp2_poly4(x1) : fast_exp2(x1);
            packed[i >> 1] = pack_half2(p0, p1);
        }
        tmax = fmaxf(tmax, tm2);
    } else {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const bool v0 = c + i < kv_valid, v1 = c + i + 1 < kv_valid;
            if (v0) tmax = fmaxf(tmax, __uint_as_float(raw[i]));
            if (v1) tmax = fmaxf(tmax, __uint_as_float(raw[i + 1]));
            const float p0 = v0 ? fast_exp2(fmaf(__uint_as_float(raw[i]), sl2, neg)) : 0.f;
            const float p1 = v1 ? fast_exp2(fmaf(__uint_as_float(raw[i + 1]), sl2, neg)) : 0.f;
            packed[i >> 1] = pack_half2(p0, p1);
        }
    }
}
__device__ __forceinline__ void store_p_chunk(uint32_t sP, int r, int c, const uint32_t (&packed)[16]) {
    const uint32_t chunk = sP + (c >> 6) * 128 * 128 + r * 128;
    const int u0 = (c & 63) >> 3;  // first 16-byte unit of this 32-column group within the 128 B row
#pragma unroll
    for (int u = 0; u < 4; ++u)
        sts128(chunk + (((u0 + u) ^ (r & 7)) << 4), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
}

template <int DPAD>
struct StreamCfg {
    static constexpr bool P_TMEM = DPAD <= 48;  // S 128 | O DPAD | P 64 must fit the 256-column allocation
};

template <int DPAD>
__global__ void __launch_bounds__(ATT_THREADS, DPAD <= 48 ? 2 : 1)
attention_stream_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
    using L = StreamSmem<DPAD>;
    constexpr int BKV = L::BKV;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint8_t* sOnes = smem + L::ONES_OFF;
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;   // [2]
    uint64_t* kv_empty = bars + 3;  // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* pv_full = bars + 7;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
    const bool ones_row = p.d < p.d16;  // column d of O doubles as the row sum
    const bool use_pt = StreamCfg<DPAD>::P_TMEM && ones_row;  // P through TMEM (the l columns are not needed then)

    uint8_t* sQ = smem;
    auto sK = [&](int stage) { return smem + L::Q_BYTES + stage * L::STAGE_BYTES; };
    auto sV = [&](int stage) { return smem + L::Q_BYTES + stage * L::STAGE_BYTES + L::K_BYTES; };
    uint8_t* sP = smem + L::P_OFF;

    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 5 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_full, 128);
        mbar_init(pv_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, L::TMEM_COLS);
    if (threadIdx.x < 128)  // 2 KiB of 1.0h (P x ones when there is no spare V^T row)
        reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    const uint32_t tmem_s = tmem_base;               // [0, 128)
    const uint32_t tmem_o = tmem_base + BKV;         // [128, 128 + DPAD)
    const uint32_t tmem_l = tmem_o + DPAD;           // [128 + DPAD, +16): P x ones (only without the ones row)
    constexpr bool PT = StreamCfg<DPAD>::P_TMEM;
    const uint32_t tmem_p = tmem_o + DPAD;           // PT: [128 + DPAD, +64): P as fp16 pairs (needs the ones row: no l columns)
    const int n_tiles = p.n_kv_tiles;

    if (warp == 4) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            mbar_expect_tx(q_full, p.nkc * 128 * 128);
            for (int kc = 0; kc < p.nkc; ++kc) tma_load_4d(sQ + kc * 128 * 128, &tmQ, q_full, kc * 64, head, q0, img);
            const uint32_t tx = p.nkc * BKV * 128 + (BKV / 64) * p.d16 * 128;
            for (int j = 0; j < n_tiles; ++j) {
                const int stage = j & 1;
                if (j >= 2) mbar_wait(&kv_empty[stage], ((j >> 1) - 1) & 1);
                mbar_expect_tx(&kv_full[stage], tx);
                for (int kc = 0; kc < p.nkc; ++kc)
                    tma_load_4d(sK(stage) + kc * BKV * 128, &tmK, &kv_full[stage], kc * 64, head, j * BKV, img);
                for (int c = 0; c < BKV / 64; ++c)
                    tma_load_4d(sV(stage) + c * L::V_CHUNK, &tmV, &kv_full[stage], j * BKV + c * 64, 0, head, img);
            }
        }
    } else if (warp == 5) {
        // ---------------------------------------------------- MMA issuer (whole warp in the loop, one lane issues)
        const int ksteps_s = (p.d + 15) / 16;
        const uint32_t qa = smem_u32(sQ), pa = smem_u32(sP), oa = smem_u32(sOnes);
        // tile j's operands have landed: plant the ones row into its V^T chunks, then S(j) = Q K(j)^T
        auto stage_ready_issue_s = [&](int j) {
            const int stage = j & 1;
            mbar_wait(&kv_full[stage], (j >> 1) & 1);
            if (ones_row && lane < (BKV / 64) * 8) {
                uint8_t* row = sV(stage) + (lane >> 3) * L::V_CHUNK + p.d * 128 + (lane & 7) * 16;
                *reinterpret_cast<uint4*>(row) = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
            }
            fence_proxy_async_smem();
            __syncwarp();
            tc_fence_after();
            if (elect_one()) {
                const uint32_t ka = smem_u32(sK(stage));
                for (int ks = 0; ks < ksteps_s; ++ks) {
                    const uint32_t off_q = (ks >> 2) * 128 * 128 + (ks & 3) * 32;
                    const uint32_t off_k = (ks >> 2) * BKV * 128 + (ks & 3) * 32;
                    umma_f16(tmem_s, umma_desc_kmajor_sw128(qa + off_q), umma_desc_kmajor_sw128(ka + off_k), p.idesc_s,
                             ks != 0 ? 1u : 0u);
                }
                umma_commit(s_full);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        stage_ready_issue_s(0);
        for (int j = 0; j < n_tiles; ++j) {
            const int stage = j & 1;
            mbar_wait(p_full, j & 1);  // P(j) is in shared memory and S(j) has been read out
            if (j + 1 < n_tiles) stage_ready_issue_s(j + 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t va = smem_u32(sV(stage));
                const uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < BKV / 16; ++ks) {
                    const uint32_t off_p = (ks >> 2) * 128 * 128 + (ks & 3) * 32;
                    const uint32_t off_v = (ks >> 2) * L::V_CHUNK + (ks & 3) * 32;
                    if (PT && use_pt)
                        umma_f16_ts(tmem_o, tmem_p + 8 * ks, umma_desc_kmajor_sw128(va + off_v), p.idesc_pv, (acc | ks) ? 1u : 0u);
                    else
                        umma_f16(tmem_o, umma_desc_kmajor_sw128(pa + off_p), umma_desc_kmajor_sw128(va + off_v), p.idesc_pv,
                                 (acc | ks) ? 1u : 0u);
                }
                if (!ones_row) {
#pragma unroll
                    for (int ks = 0; ks < BKV / 16; ++ks) {
                        const uint32_t off_p = (ks >> 2) * 128 * 128 + (ks & 3) * 32;
                        umma_f16(tmem_l, umma_desc_kmajor_sw128(pa + off_p), umma_desc_kmajor_sw128(oa), p.idesc_l,
                                 (acc | ks) ? 1u : 0u);
                    }
                }
                umma_commit(pv_full);
                umma_commit(&kv_empty[stage]);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------- softmax + epilogue: thread == query row
        const int r = warp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const float sl2 = p.scale_log2e;
        const uint32_t sP_a = smem_u32(sP);
        auto put_p = [&](int c, const uint32_t (&packed)[16]) {
            if (PT && use_pt) tmem_st_32x16(tmem_p + lane_off + (c >> 1), packed);
            else store_p_chunk(sP_a, r, c, packed);
        };
        float mr = -INFINITY;  // lazy reference maximum, already multiplied by scale * log2(e)
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int kv_valid = min(BKV, p.Nk - j * BKV);  // columns >= kv_valid are TMA zero fill: masked out
            const bool full_tile = kv_valid == BKV;         // warp-uniform
            float tmax = -INFINITY;
            bool redo = (j == 0);
            if (!redo) {
                // ---- fast path: exponentials against the stale reference, one TMEM read
                const float neg = -mr;
                uint32_t ra[32], rb[32], packed[16];
                tmem_ld_32x32(tmem_s + lane_off, ra);
                tmem_ld_32x32(tmem_s + lane_off + 32, rb);
                tmem_ld_wait();
                if (full_tile) softmax_chunk<false>(ra, sl2, neg, 0, kv_valid, packed, tmax);
                else softmax_chunk<true>(ra, sl2, neg, 0, kv_valid, packed, tmax);
                tmem_ld_32x32(tmem_s + lane_off + 64, ra);
                // P V(j-1) reads the P buffer and accumulates into O: both must be done before P(j) is written
                mbar_wait(pv_full, (j - 1) & 1);
                put_p(0, packed);
                if (full_tile) softmax_chunk<false>(rb, sl2, neg, 32, kv_valid, packed, tmax);
                else softmax_chunk<true>(rb, sl2, neg, 32, kv_valid, packed, tmax);
                put_p(32, packed);
                tmem_ld_wait();
                tmem_ld_32x32(tmem_s + lane_off + 96, rb);
                if (full_tile) softmax_chunk<false>(ra, sl2, neg, 64, kv_valid, packed, tmax);
                else softmax_chunk<true>(ra, sl2, neg, 64, kv_valid, packed, tmax);
                put_p(64, packed);
                tmem_ld_wait();
                if (full_tile) softmax_chunk<false>(rb, sl2, neg, 96, kv_valid, packed, tmax);
                else softmax_chunk<true>(rb, sl2, neg, 96, kv_valid, packed, tmax);
                put_p(96, packed);
                redo = __any_sync(0xffffffffu, tmax * sl2 > mr + ATT_TAU);
            } else {
                // first tile: the row maximum is needed before anything else
#pragma unroll 1
                for (int c = 0; c < BKV; c += 32) {
                    uint32_t raw[32];
                    tmem_ld_32x32(tmem_s + lane_off + c, raw);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c + i < kv_valid) tmax = fmaxf(tmax, __uint_as_float(raw[i]));
                }
            }
            if (redo) {
                // ---- slow path (warp-uniform): move the reference of the violating rows, rescale O / l, redo this tile's P
                const float tm = tmax * sl2;
                const bool fix = tm > mr + ATT_TAU;            // first tile: mr = -inf -> every row
                const float mr_new = fix ? tm : mr;
                const float alpha = fix ? fast_exp2(mr - mr_new) : 1.0f;
                mr = mr_new;
                if (j > 0) {
                    tc_fence_after();  // pv_full(j-1) was waited for above: O / l are quiescent
                    const int ncols = ones_row ? DPAD : DPAD + 16;  // without the ones row, l sits right after O
#pragma unroll 1
                    for (int c = 0; c < ncols; c += 16) {
                        uint32_t raw[16];
                        tmem_ld_32x16(tmem_o + lane_off + c, raw);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * alpha);
                        tmem_st_32x16(tmem_o + lane_off + c, raw);
                    }
                    tmem_st_wait();
                }
                const float neg = -mr;
                float dummy = 0.f;
#pragma unroll 1
                for (int c = 0; c < BKV; c += 32) {
                    uint32_t raw[32], packed[16];
                    tmem_ld_32x32(tmem_s + lane_off + c, raw);
                    tmem_ld_wait();
                    softmax_chunk<true>(raw, sl2, neg, c, kv_valid, packed, dummy);
                    put_p(c, packed);
                }
            }
            if (PT && use_pt) tmem_st_wait();
            else fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // ---- final: normalise, store
        mbar_wait(pv_full, (n_tiles - 1) & 1);
        tc_fence_after();
        uint32_t lfin;
        tmem_ld_32x1((ones_row ? tmem_o + p.d : tmem_l) + lane_off, lfin);
        tmem_ld_wait();
        const float l_tot = __uint_as_float(lfin);
        const float inv_l = 1.0f / l_tot;
        const bool row_ok = (q0 + r) < p.Nq;
        if (p.lse && row_ok)
            p.lse[(static_cast<long long>(img) * p.heads + head) * p.Nq + q0 + r] = mr + log2f(l_tot);
        __half* orow = p.out + (static_cast<long long>(img) * p.Nq + q0 + r) * p.ldo + head * p.d;
#pragma unroll
        for (int c = 0; c < DPAD; c += 16) {
            if (c < p.d) {  // warp-uniform
                uint32_t raw[16];
                tmem_ld_32x16(tmem_o + lane_off + c, raw);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        if (c + g * 8 < p.d) {
                            uint4 u;
                            __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                h[e] = __floats2half2_rn(__uint_as_float(raw[g * 8 + 2 * e]) * inv_l,
                                                         __uint_as_float(raw[g * 8 + 2 * e + 1]) * inv_l);
                            *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

// =====================================================================================================================
// d_head 40 (SD1.5's 64x64 / 32x32 self-attention: d < 48 leaves the spare V^T row for the row sums): 64-key tiles,
// S double-buffered in TMEM, eight softmax warps on split keys.
// ncu on attention_stream_kernel<48> (profiles/prof_attn_stream_r1j) showed the four softmax warps waiting a quarter of
// their time for S(j): the MMA warp can only start S(j+1) once softmax(j) has read the single S buffer.  Here
//   * the key tile is 64 wide and S is DOUBLE-BUFFERED (S(j+1), S(j+2) are issued while softmax(j) runs; a row thread finds
//     its logits waiting); K/V ride a four-stage ring;
//   * warps w and w + 4 own the same 32 TMEM lanes (query rows) and run INDEPENDENT online-softmax streams over the lower /
//     upper 32 keys of every tile -- own lazy reference, own accumulator (O_A, O_B), own row sum -- merged once at the end
//     like split-KV decoding:  O = (2^(mA-m) O_A + 2^(mB-m) O_B) / (2^(mA-m) l_A + 2^(mB-m) l_B),  m = max(mA, mB).
//     No per-tile exchange between the two warps of a row; four row-math warps per scheduler instead of two.
//   TMEM: S0 64 | S1 64 | O_A 48 | O_B 48 | P 32 = 256 columns, two CTAs per SM.
// Measured at 8 img x 8 heads x 4096 x 4096, d = 40 (CUDA events, tools/time_attn.py; MUFU bound 239 us):
//   single S buffer, 128 keys (kernel above)                     434 us   XU pipe 57 %
//   double-buffered S, 64 keys, 4 softmax warps                  383 us   XU pipe 65 % (profiles/prof_attn_stream64_r2w)
//   + 8 warps, row maxima exchanged per tile (smem + bar.sync)   407 us   the exchange costs more than the warps hide
//   + 8 warps, independent key halves (this kernel)              374 us
//   exp2 on the FMA pipe for 1/6, 1/4, 1/2 of the elements       384 / 391 / 434 us (on the 4-warp form): not XU-throughput
//   bound -- the warps stall on fixed-latency dependencies (ncu: stall_wait 31 %) -- so none of it is kept;
//   prefetching S(j+1) from TMEM before handing over P(j)        411 us (4 warps), 788 us (8 warps: spills): dropped.
struct Stream64 {
    static constexpr int BKV = 64;
    static constexpr int DPAD = 48;
    static constexpr int NSTAGE = 4;
    static constexpr int Q_BYTES = 128 * 128;
    static constexpr int K_BYTES = BKV * 128;
    static constexpr int V_BYTES = DPAD * 128;
    static constexpr int STAGE_BYTES = K_BYTES + V_BYTES;
    static constexpr int BAR_OFF = Q_BYTES + NSTAGE * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static constexpr int TMEM_COLS = 256;
};

constexpr int ATT64S_THREADS = 320;  // warps 0-7 softmax (lane quarter = warp & 3, key half = warp >> 2), 8 TMA, 9 MMA

__global__ void __launch_bounds__(ATT64S_THREADS, 2)
attention_stream64s_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p) {
    using L = Stream64;
    constexpr int BKV = L::BKV, DPAD = L::DPAD, NS = L::NSTAGE;
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;        // [NS]
    uint64_t* kv_empty = bars + 1 + NS;  // [NS]
    uint64_t* s_full = bars + 1 + 2 * NS;  // [2]
    uint64_t* p_full = s_full + 2;
    uint64_t* pv_full = s_full + 3;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_full + 4);
    float* xm = reinterpret_cast<float*>(smem + L::BAR_OFF + 256);  // [128]: the upper half's final reference (in the 1 KiB tail)

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128, head = blockIdx.y, img = blockIdx.z;
    uint8_t* sQ = smem;
    auto sK = [&](int stage) { return smem + L::Q_BYTES + stage * L::STAGE_BYTES; };
    auto sV = [&](int stage) { return smem + L::Q_BYTES + stage * L::STAGE_BYTES + L::K_BYTES; };

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 9 && lane == 0) {
        mbar_init(q_full, 1);
        for (int i = 0; i < NS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        mbar_init(p_full, 256);
        mbar_init(pv_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, L::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    const uint32_t tmem_s = tmem_base;                 // S0 [0, 64) | S1 [64, 128)
    const uint32_t tmem_o = tmem_base + 2 * BKV;       // O_A [128, 176) | O_B [176, 224): column d is the row sum
    const uint32_t tmem_p = tmem_o + 2 * DPAD;         // [224, 256): P as fp16 pairs, lower | upper key half
    const int n_tiles = p.n_kv_tiles;

    if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer
            mbar_expect_tx(q_full, 128 * 128);
            tma_load_4d(sQ, &tmQ, q_full, 0, head, q0, img);
            const uint32_t tx = BKV * 128 + p.d16 * 128;
            for (int j = 0; j < n_tiles; ++j) {
                const int stage = j % NS;
                if (j >= NS) mbar_wait(&kv_empty[stage], ((j / NS) - 1) & 1);
                mbar_expect_tx(&kv_full[stage], tx);
                tma_load_4d(sK(stage), &tmK, &kv_full[stage], 0, head, j * BKV, img);
                tma_load_4d(sV(stage), &tmV, &kv_full[stage], j * BKV, 0, head, img);
            }
        }
    } else if (warp == 9) {
        // ---------------------------------------------------- MMA issuer (whole warp in the loop, one lane issues)
        const int ksteps_s = (p.d + 15) / 16;
        const uint32_t qa = smem_u32(sQ);
        auto issue_s = [&](int j) {
            const int stage = j % NS;
            mbar_wait(&kv_full[stage], (j / NS) & 1);
            if (lane < 8)
                *reinterpret_cast<uint4*>(sV(stage) + p.d * 128 + lane * 16) =
                    make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
            fence_proxy_async_smem();
            __syncwarp();
            tc_fence_after();
            if (elect_one()) {
                const uint32_t ka = smem_u32(sK(stage));
                for (int ks = 0; ks < ksteps_s; ++ks)
                    umma_f16(tmem_s + (j & 1) * BKV, umma_desc_kmajor_sw128(qa + ks * 32), umma_desc_kmajor_sw128(ka + ks * 32),
                             p.idesc_s, ks != 0 ? 1u : 0u);
                umma_commit(&s_full[j & 1]);
            }
            __syncwarp();
        };
        mbar_wait(q_full, 0);
        issue_s(0);
        if (n_tiles > 1) issue_s(1);
        for (int j = 0; j < n_tiles; ++j) {
            const int stage = j % NS;
            mbar_wait(p_full, j & 1);  // P(j) is in TMEM and S(j) has been read out
            tc_fence_after();
            if (elect_one()) {
                const uint32_t va = smem_u32(sV(stage));
                const uint32_t acc = j > 0 ? 1u : 0u;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        umma_f16_ts(tmem_o + h * DPAD, tmem_p + h * 16 + 8 * ks, umma_desc_kmajor_sw128(va + (2 * h + ks) * 32),
                                    p.idesc_pv, (acc | ks) ? 1u : 0u);
                umma_commit(pv_full);
                umma_commit(&kv_empty[stage]);
            }
            __syncwarp();
            if (j + 2 < n_tiles) issue_s(j + 2);
        }
    } else {
        // ---------------------------------------------------- softmax: thread == (query row, key half)
        const int quarter = warp & 3, half = warp >> 2;
        const int r = quarter * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
        const float sl2 = p.scale_log2e;
        const uint32_t tp = tmem_p + lane_off + half * 16;
        const uint32_t to = tmem_o + lane_off + half * DPAD;
        float mr = -INFINITY;  // this half's lazy reference maximum, already multiplied by scale * log2(e)
        for (int j = 0; j < n_tiles; ++j) {
            const uint32_t ts = tmem_s + lane_off + (j & 1) * BKV + half * 32;
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const int kv_valid = max(0, min(32, p.Nk - j * BKV - half * 32));  // columns >= Nk are TMA zero fill
            const bool full_tile = kv_valid == 32;  // warp-uniform
            float tmax = -INFINITY;
            uint32_t raw[32], packed[16];
            tmem_ld_32x32(ts, raw);
            tmem_ld_wait();
            bool redo = (j == 0);
            if (!redo) {
                if (full_tile) softmax_chunk<false, 0>(raw, sl2, -mr, 0, kv_valid, packed, tmax);
                else softmax_chunk<true, 0>(raw, sl2, -mr, 0, kv_valid, packed, tmax);
                // P V(j-1) reads the P columns and accumulates into O: it must be done before P(j) is written
                mbar_wait(pv_full, (j - 1) & 1);
                tmem_st_32x16(tp, packed);
                redo = __any_sync(0xffffffffu, tmax * sl2 > mr + ATT_TAU);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i < kv_valid) tmax = fmaxf(tmax, __uint_as_float(raw[i]));
            }
            if (redo) {
                const float tm = tmax * sl2;
                const bool fix = tm > mr + ATT_TAU;            // first tile: mr = -inf -> every row
                const float mr_new = fix ? tm : mr;
                const float alpha = fix ? fast_exp2(mr - mr_new) : 1.0f;
                mr = mr_new;
                if (j > 0) {
                    tc_fence_after();  // pv_full(j-1) was waited for above: O is quiescent
#pragma unroll 1
                    for (int c = 0; c < DPAD; c += 16) {
                        uint32_t o[16];
                        tmem_ld_32x16(to + c, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                        tmem_st_32x16(to + c, o);
                    }
                }
                float dummy = 0.f;
                softmax_chunk<true, 0>(raw, sl2, -mr, 0, kv_valid, packed, dummy);
                tmem_st_32x16(tp, packed);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(p_full);
        }
        // ---- merge the two key halves, normalise, store (the lower-half warps)
        if (half == 1) xm[r] = mr;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
        if (half == 0) {
            const float mb = xm[r];
            const float m = fmaxf(mr, mb);
            const float wa = fast_exp2(mr - m), wb = fast_exp2(mb - m);
            mbar_wait(pv_full, (n_tiles - 1) & 1);
            tc_fence_after();
            uint32_t la, lb;
            tmem_ld_32x1(tmem_o + lane_off + p.d, la);
            tmem_ld_32x1(tmem_o + lane_off + DPAD + p.d, lb);
            tmem_ld_wait();
            const float l_tot = wa * __uint_as_float(la) + wb * __uint_as_float(lb);
            const float inv_l = 1.0f / l_tot;
            const float ca = wa * inv_l, cb = wb * inv_l;
            const bool row_ok = (q0 + r) < p.Nq;
            if (p.lse && row_ok)
                p.lse[(static_cast<long long>(img) * p.heads + head) * p.Nq + q0 + r] = m + log2f(l_tot);
            __half* orow = p.out + (static_cast<long long>(img) * p.Nq + q0 + r) * p.ldo + head * p.d;
#pragma unroll
            for (int c = 0; c < DPAD; c += 16) {
                if (c < p.d) {  // warp-uniform
                    uint32_t oa[16], ob[16];
                    tmem_ld_32x16(tmem_o + lane_off + c, oa);
                    tmem_ld_32x16(tmem_o + lane_off + DPAD + c, ob);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            if (c + g * 8 < p.d) {
                                uint4 u;
                                __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const int i0 = g * 8 + 2 * e;
                                    h[e] = __floats2half2_rn(
                                        fmaf(ca, __uint_as_float(oa[i0]), cb * __uint_as_float(ob[i0])),
                                        fmaf(ca, __uint_as_float(oa[i0 + 1]), cb * __uint_as_float(ob[i0 + 1])));
                                }
                                *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
                            }
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

static int launch_attn_stream64s(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                                 dim3 grid, cudaStream_t stream) {
    using L = Stream64;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attention_stream64s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) != cudaSuccess)
            return CTRLORA_ERR_CUDA;
        attr = true;
    }
    if (launch_pdl(attention_stream64s_kernel, grid, dim3(ATT64S_THREADS), (size_t)L::TOTAL, stream, tq, tk, tv, p) != cudaSuccess)
        return CTRLORA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

// =====================================================================================================================
// Cross-attention to a short context (Nk <= 128: the 77 text tokens), d_head <= 80: the 64x64 and 32x32 levels.
// attention_kernel<.., false> gives every (128 queries, head, image) its own CTA, and each CTA is one short serial chain
// (TMA -> S -> two TMEM passes over 128 columns -> P -> P V over 128 keys -> epilogue): 46 us for 3 GF at the 64x64 level.
// This kernel is PERSISTENT over the query tiles of one (head, image): K and V^T are loaded once, Q tiles ride a two-stage
// TMA ring, S and O are double-buffered in TMEM, and the softmax warps run the epilogue of tile i-1 AFTER the softmax of
// tile i, so the S / P V products and their latencies hide behind row math.  (Four dedicated epilogue warps were tried
// instead: 36 us against 29 us at the 64x64 level -- dropped.)  Measured (tools/time_attn_cross.py, batch 8): 64x64 level
// 43.8 -> 29.0 us, 32x32 level 16.7 -> 12.7 us; batch 16: 77 -> 52 us.  Only ceil(Nk / 32) 32-column chunks of S are
// read (once: the tile's logits stay in registers between the max and the exp pass) and P V runs ceil(Nk / 16) k-steps.
template <int DPAD>
struct CrossSmem {
    static constexpr int NKC = (DPAD + 63) / 64;
    static constexpr int K_BYTES = NKC * 128 * 128;
    static constexpr int V_CHUNK = DPAD * 128;
    static constexpr int V_BYTES = 2 * V_CHUNK;
    static constexpr int Q_STAGE = NKC * 128 * 128;
    static constexpr int P_BUF = 2 * 128 * 128;
    static constexpr int K_OFF = 0;
    static constexpr int V_OFF = K_BYTES;
    static constexpr int Q_OFF = V_OFF + V_BYTES + ((1024 - (V_BYTES & 1023)) & 1023);  // 1 KiB aligned (SWIZZLE_128B)
    static constexpr int P_OFF = Q_OFF + 2 * Q_STAGE;
    static constexpr int ONES_OFF = P_OFF + 2 * P_BUF;
    static constexpr int BAR_OFF = ONES_OFF + 2048;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;
    static constexpr int O_STRIDE = DPAD + 16;  // O | l (row sums from P x ones)
    static constexpr int TMEM_COLS = 512;       // S0 128 | S1 128 | (O | l) x 2
};

template <int DPAD>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_cross_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ AttnParams p, int splits) {
    using L = CrossSmem<DPAD>;
    pdl_launch_dependents();
    // this CTA's query tiles of one (head, image)
    const int pair = blockIdx.x / splits, part = blockIdx.x % splits;
    const int head = pair % p.heads, img = pair / p.heads;
    const int n_q_tiles = (p.Nq + 127) / 128;
    const int t0 = static_cast<int>(static_cast<long long>(n_q_tiles) * part / splits);
    const int t1 = static_cast<int>(static_cast<long long>(n_q_tiles) * (part + 1) / splits);
    const int n = t1 - t0;
    if (n <= 0) return;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* kv_full = bars;
    uint64_t* q_full = bars + 1;    // [2]
    uint64_t* q_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;    // [2]
    uint64_t* p_full = bars + 7;    // [2]
    uint64_t* pv_full = bars + 9;   // [2]
    uint64_t* o_free = bars + 11;   // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);
    uint8_t* sK = smem + L::K_OFF;
    uint8_t* sV = smem + L::V_OFF;
    uint8_t* sOnes = smem + L::ONES_OFF;
    auto sQ = [&](int st) { return smem + L::Q_OFF + st * L::Q_STAGE; };
    auto sP = [&](int b) { return smem + L::P_OFF + b * L::P_BUF; };

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    if (warp == 4 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
    }
    if (warp == 5 && lane == 0) {
        mbar_init(kv_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 128);
            mbar_init(&pv_full[i], 1);
            mbar_init(&o_free[i], 128);
        }
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, L::TMEM_COLS);
    if (threadIdx.x < 128)  // 128 x 16 B = 2 KiB of 1.0h; all elements equal, so the swizzle is irrelevant
        reinterpret_cast<uint4*>(sOnes)[threadIdx.x] = make_uint4(0x3C003C00u, 0x3C003C00u, 0x3C003C00u, 0x3C003C00u);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();
    const uint32_t tmem_s = tmem_base;  // S buffer b at + 128 b
    auto tmem_o = [&](int b) { return tmem_base + 256 + b * L::O_STRIDE; };
    const int ksteps_pv = (p.Nk + 15) / 16;

    if (warp == 4) {
        if (lane == 0) {
            // ------------------------------------------------ TMA producer: K, V^T once, then the Q tiles
            mbar_expect_tx(kv_full, p.nkc * 128 * 128 + 2 * p.d16 * 128);
            for (int kc = 0; kc < p.nkc; ++kc) tma_load_4d(sK + kc * 128 * 128, &tmK, kv_full, kc * 64, head, 0, img);
            for (int c = 0; c < 2; ++c) tma_load_4d(sV + c * L::V_CHUNK, &tmV, kv_full, c * 64, 0, head, img);
            for (int i = 0; i < n; ++i) {
                const int st = i & 1;
                if (i >= 2) mbar_wait(&q_empty[st], ((i >> 1) - 1) & 1);
                mbar_expect_tx(&q_full[st], p.nkc * 128 * 128);
                for (int kc = 0; kc < p.nkc; ++kc)
                    tma_load_4d(sQ(st) + kc * 128 * 128, &tmQ, &q_full[st], kc * 64, head, (t0 + i) * 128, img);
            }
        }
    } else if (warp == 5) {
        // ---------------------------------------------------- MMA issuer (whole warp in the loop, one lane issues)
        const int ksteps_s = (p.d + 15) / 16;
        const uint32_t ka = smem_u32(sK), va = smem_u32(sV), oa = smem_u32(sOnes);
        auto issue_s = [&](int i) {
            const int st = i & 1;
            mbar_wait(&q_full[st], (i >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t qa = smem_u32(sQ(st));
                for (int ks = 0; ks < ksteps_s; ++ks) {
                    const uint32_t off = (ks >> 2) * 128 * 128 + (ks & 3) * 32;
                    umma_f16(tmem_s + st * 128, umma_desc_kmajor_sw128(qa + off), umma_desc_kmajor_sw128(ka + off), p.idesc_s,
                             ks != 0 ? 1u : 0u);
                }
                umma_commit(&s_full[st]);
                umma_commit(&q_empty[st]);
            }
            __syncwarp();
        };
        mbar_wait(kv_full, 0);
        issue_s(0);
        for (int i = 0; i < n; ++i) {
            const int b = i & 1;
            if (i + 1 < n) issue_s(i + 1);  // its S buffer was read out before p_full(i-1), waited for last iteration
            mbar_wait(&p_full[b], (i >> 1) & 1);
            if (i >= 2) mbar_wait(&o_free[b], ((i >> 1) - 1) & 1);  // the epilogue of tile i-2 has read this O buffer
            tc_fence_after();
            if (elect_one()) {
                const uint32_t pa = smem_u32(sP(b));
                for (int ks = 0; ks < ksteps_pv; ++ks) {
                    const uint32_t off_p = (ks >> 2) * 128 * 128 + (ks & 3) * 32;
                    const uint32_t off_v = (ks >> 2) * L::V_CHUNK + (ks & 3) * 32;
                    umma_f16(tmem_o(b), umma_desc_kmajor_sw128(pa + off_p), umma_desc_kmajor_sw128(va + off_v), p.idesc_pv,
                             ks != 0 ? 1u : 0u);
                }
                for (int ks = 0; ks < ksteps_pv; ++ks) {
                    const uint32_t off_p = (ks >> 2) * 128 * 128 + (ks & 3) * 32;
                    umma_f16(tmem_o(b) + DPAD, umma_desc_kmajor_sw128(pa + off_p), umma_desc_kmajor_sw128(oa), p.idesc_l,
                             ks != 0 ? 1u : 0u);
                }
                umma_commit(&pv_full[b]);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------- softmax + epilogue: thread == query row
        const int r = warp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        const float sl2 = p.scale_log2e;
        const int nch = (p.Nk + 31) / 32;  // 32-column chunks that hold keys (warp-uniform)
        float m_tile[2] = {0.f, 0.f};
        auto epilogue = [&](int j) {
            const int b = j & 1;
            mbar_wait(&pv_full[b], (j >> 1) & 1);
            tc_fence_after();
            uint32_t lraw;
            tmem_ld_32x1(tmem_o(b) + DPAD + lane_off, lraw);
            tmem_ld_wait();
            const float l_tot = __uint_as_float(lraw);
            const float inv_l = 1.0f / l_tot;
            const int row = (t0 + j) * 128 + r;
            const bool row_ok = row < p.Nq;
            if (p.lse && row_ok) p.lse[(static_cast<long long>(img) * p.heads + head) * p.Nq + row] = m_tile[b] * sl2 + log2f(l_tot);
            __half* orow = p.out + (static_cast<long long>(img) * p.Nq + row) * p.ldo + head * p.d;
#pragma unroll
            for (int c = 0; c < DPAD; c += 16) {
                if (c < p.d) {  // warp-uniform
                    uint32_t raw[16];
                    tmem_ld_32x16(tmem_o(b) + lane_off + c, raw);
                    tmem_ld_wait();
                    if (row_ok) {
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            if (c + g * 8 < p.d) {
                                uint4 u;
                                __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
                                for (int e = 0; e < 4; ++e)
                                    h[e] = __floats2half2_rn(__uint_as_float(raw[g * 8 + 2 * e]) * inv_l,
                                                             __uint_as_float(raw[g * 8 + 2 * e + 1]) * inv_l);
                                *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&o_free[b]);
        };
        for (int i = 0; i < n; ++i) {
            const int b = i & 1;
            mbar_wait(&s_full[b], (i >> 1) & 1);
            tc_fence_after();
            // the tile's logits, read once
            uint32_t s[4][32];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < nch) tmem_ld_32x32(tmem_s + b * 128 + lane_off + 32 * k, s[k]);
            tmem_ld_wait();
            float m = -INFINITY;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < nch) {
#pragma unroll
                    for (int c = 0; c < 32; ++c)
                        if (32 * k + c < p.Nk) m = fmaxf(m, __uint_as_float(s[k][c]));
                }
            }
            m_tile[b] = m;
            const float neg = -m * sl2;
            const uint32_t pbase = smem_u32(sP(b));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (k < nch) {
                    uint32_t packed[16];
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
                        const float p0 = (32 * k + c < p.Nk) ? fast_exp2(fmaf(__uint_as_float(s[k][c]), sl2, neg)) : 0.f;
                        const float p1 = (32 * k + c + 1 < p.Nk) ? fast_exp2(fmaf(__uint_as_float(s[k][c + 1]), sl2, neg)) : 0.f;
                        packed[c >> 1] = pack_half2(p0, p1);
                    }
                    store_p_chunk(pbase, r, 32 * k, packed);
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&p_full[b]);
            if (i >= 1) epilogue(i - 1);  // P V(i-1) has long finished: its latency hid behind this tile's row math
        }
        epilogue(n - 1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem_base, L::TMEM_COLS);
    }
}

template <int DPAD>
static int launch_attn_cross(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                             int batch, cudaStream_t stream) {
    using L = CrossSmem<DPAD>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attention_cross_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) !=
            cudaSuccess)
            return CTRLORA_ERR_CUDA;
        attr = true;
    }
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const int pairs = p.heads * batch, n_q_tiles = (p.Nq + 127) / 128;
    int splits = sms / pairs;  // (head, image) pairs first; split their query tiles only while SMs are left over
    if (splits < 1) splits = 1;
    if (splits > n_q_tiles) splits = n_q_tiles;
    if (launch_pdl(attention_cross_kernel<DPAD>, dim3(pairs * splits), dim3(ATT_THREADS), (size_t)L::TOTAL, stream, tq, tk, tv, p,
                   splits) != cudaSuccess)
        return CTRLORA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

template <int DPAD>
static int launch_attn_stream(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                              dim3 grid, cudaStream_t stream) {
    using L = StreamSmem<DPAD>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attention_stream_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL) !=
            cudaSuccess)
            return CTRLORA_ERR_CUDA;
        attr = true;
    }
    if (launch_pdl(attention_stream_kernel<DPAD>, grid, dim3(ATT_THREADS), (size_t)L::TOTAL, stream, tq, tk, tv, p) !=
        cudaSuccess)
        return CTRLORA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

template <int DPAD, int BKV, bool MULTI>
static int launch_attn(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p, dim3 grid,
                       cudaStream_t stream) {
    using L = AttnSmem<DPAD, BKV, MULTI>;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(attention_kernel<DPAD, BKV, MULTI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 L::TOTAL) != cudaSuccess)
            return CTRLORA_ERR_CUDA;
        attr = true;
    }
    if (launch_pdl(attention_kernel<DPAD, BKV, MULTI>, grid, dim3(ATT_THREADS), (size_t)L::TOTAL, stream, tq, tk, tv, p) !=
        cudaSuccess)
        return CTRLORA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}

}  // namespace ctrl

using namespace ctrl;

extern "C" int ctrlora_attention_f16(const void* q, long long ldq, const void* k, long long ldk, const void* vt,
                                     int nk_pad, void* out, long long ldo, float* lse, int batch, int heads, int nq,
                                     int nk, int head_dim, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!q || !k || !vt || !out) return CTRLORA_ERR_ARG;
    const int d = head_dim;
    if (d % 8 != 0 || d > 160 || nk_pad % 8 != 0 || nk_pad < nk || ldq % 8 != 0 || ldk % 8 != 0 || ldo % 8 != 0)
        return CTRLORA_ERR_ARG;
    AttnParams p;
    memset(&p, 0, sizeof(p));
    p.Nq = nq; p.Nk = nk; p.heads = heads; p.d = d; p.d16 = (d + 15) / 16 * 16; p.nkc = (d + 63) / 64;
    p.scale_log2e = (1.0f / sqrtf(static_cast<float>(d))) * 1.4426950408889634f;
    p.out = reinterpret_cast<__half*>(out); p.ldo = ldo; p.lse = lse;
    const bool multi = nk > 256;
    if (multi && d > 80) return CTRLORA_ERR_UNSUPPORTED;  // d_head 160 with > 256 keys: not on the 512x512 path
    static int stream64_env = -1;
    if (stream64_env < 0) {
        const char* e = getenv("CTRLORA_ATTN_STREAM64");
        stream64_env = (e && e[0] == '0') ? 0 : 1;  // 0: the 128-key single-S-buffer kernel (kept for A/B measurements)
    }
    const bool s64 = multi && stream64_env && d < 48 && (d % 16) != 0;  // needs the spare V^T row and d16 <= 48
    const int bkv = s64 ? 64 : multi ? 128 : (nk <= 128 ? 128 : 256);
    p.n_kv_tiles = (nk + bkv - 1) / bkv;
    p.idesc_s = umma_idesc_f16(128, bkv, 0);
    p.idesc_pv = umma_idesc_f16(128, p.d16, 0);
    p.idesc_l = umma_idesc_f16(128, 16, 0);
    CUtensorMap tq, tk, tv;
    {
        uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)nq, (uint64_t)batch};
        uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ldq * 2, (uint64_t)ldq * 2 * nq};
        uint32_t box[4] = {64, 1, 128, 1};
        int rc = make_tmap_f16(&tq, q, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)d, (uint64_t)heads, (uint64_t)nk, (uint64_t)batch};
        uint64_t str[3] = {(uint64_t)d * 2, (uint64_t)ldk * 2, (uint64_t)ldk * 2 * nk};
        uint32_t box[4] = {64, 1, (uint32_t)bkv, 1};
        int rc = make_tmap_f16(&tk, k, 4, dims, str, box);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)nk, (uint64_t)d, (uint64_t)heads, (uint64_t)batch};
        uint64_t str[3] = {(uint64_t)nk_pad * 2, (uint64_t)nk_pad * 2 * d, (uint64_t)nk_pad * 2 * d * heads};
        uint32_t box[4] = {64, (uint32_t)p.d16, 1, 1};
        int rc = make_tmap_f16(&tv, vt, 4, dims, str, box);
        if (rc) return rc;
    }
    dim3 grid((nq + 127) / 128, heads, batch);
    if (s64) return launch_attn_stream64s(tq, tk, tv, p, grid, stream);
    if (multi) {
        static int stream_env = -1;
        if (stream_env < 0) {
            const char* e = getenv("CTRLORA_ATTN_STREAM");
            stream_env = (e && e[0] == '0') ? 0 : 1;  // 0: the older per-tile-fold kernel (kept for A/B measurements)
        }
        if (stream_env) {
            if (d <= 48) return launch_attn_stream<48>(tq, tk, tv, p, grid, stream);
            return launch_attn_stream<80>(tq, tk, tv, p, grid, stream);
        }
        if (d <= 48) return launch_attn<48, 128, true>(tq, tk, tv, p, grid, stream);
        return launch_attn<80, 128, true>(tq, tk, tv, p, grid, stream);
    }
    if (bkv == 128 && d <= 80) {
        static int cross_env = -1;
        if (cross_env < 0) {
            const char* e = getenv("CTRLORA_ATTN_CROSS");
            cross_env = (e && e[0] == '0') ? 0 : 1;  // 0: one CTA per query tile (kept for A/B measurements)
        }
        if (cross_env) {
            AttnParams pc = p;
            pc.idesc_s = umma_idesc_f16(128, (nk + 15) / 16 * 16, 0);  // only the key columns that exist
            if (d <= 48) return launch_attn_cross<48>(tq, tk, tv, pc, batch, stream);
            return launch_attn_cross<80>(tq, tk, tv, pc, batch, stream);
        }
    }
    if (bkv == 128) {
        if (d <= 48) return launch_attn<48, 128, false>(tq, tk, tv, p, grid, stream);
        if (d <= 80) return launch_attn<80, 128, false>(tq, tk, tv, p, grid, stream);
        return launch_attn<160, 128, false>(tq, tk, tv, p, grid, stream);
    }
    if (d <= 48) return launch_attn<48, 256, false>(tq, tk, tv, p, grid, stream);
    if (d <= 80) return launch_attn<80, 256, false>(tq, tk, tv, p, grid, stream);
    return launch_attn<160, 256, false>(tq, tk, tv, p, grid, stream);
}
