// Weight-gradient GEMM ("TN"): C[P, Q] = sum_m A[m, P] * B[m, Q]   (fp16 operands, fp32 result)
// for the trainable set of the CtrLoRA finetune (reference optimizer filter cldm/cldm_ctrlora_finetune.py:88-100):
//   LoRA   dUp   = dY^T (X Down^T),  dDown = (dY Up)^T X      (factored: no dense dW is ever formed)
//   zero-conv dW = dY^T H            (1x1 convs, cldm/cldm.py:281-282)
// Both operands are row-major over the token dimension m, i.e. "MN-major" for the tensor core: TMA boxes of
// [64 tokens][64 features] (128-byte rows, SWIZZLE_128B) are consumed by tcgen05.mma with a_major = b_major = MN.
// The token dimension is split across CTAs; every CTA parks its fp32 partial tile in a workspace slice and a second
// kernel sums the slices in a fixed order (deterministic; no atomics).
#include "common.cuh"
#include "ctrlora_b200.h"
#include "gemm_sm100.cuh"

namespace ctrl {

int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

constexpr int WG_BK = 64;          // tokens per pipeline stage
constexpr int WG_STAGES = 8;        // upper bound; the launcher fits as many as 200 KiB allow
constexpr int WG_THREADS = 192;    // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
constexpr int WG_A_BYTES = 2 * WG_BK * 128;  // two 64-feature atoms of the 128-row P tile

struct WgradParams {
    int P, Q, M;
    int q_tile;        // multiple of 64, <= 256
    int q_atoms;       // q_tile / 64
    int p_tiles, q_tiles, splits;
    int kiters_per_split;
    int stage_bytes, stages;
    uint32_t idesc;
    float* ws;         // [splits][P_pad][Q_pad] fp32, P_pad = p_tiles*128, Q_pad = q_tiles*q_tile
    // direct mode (splits == 1: enough output tiles to fill the machine, e.g. the dense conv gradients of pretraining):
    // the epilogue applies alpha / beta itself and writes the result row-major -- no workspace round trip, no reduce kernel
    int direct;
    float* out;
    long long ldo;
    float alpha, beta;
};

// MN-major operand, SWIZZLE_128B: 64-feature atoms (128 B rows), 8-token groups 1024 B apart (SBO), atoms `lbo` apart.
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ WgradParams p) {
    pdl_launch_dependents();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + WG_STAGES;
    uint64_t* tfull = bars + 2 * WG_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();

    const int tile = blockIdx.x;
    const int pt = tile % p.p_tiles, qt = tile / p.p_tiles;
    const int split = blockIdx.y;
    const int k_total = (p.M + WG_BK - 1) / WG_BK;
    const int it0 = split * p.kiters_per_split, it1 = min(k_total, it0 + p.kiters_per_split);
    const int p0 = pt * 128, q0 = qt * p.q_tile;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = WG_A_BYTES + p.q_atoms * WG_BK * 128;
            for (int it = it0; it < it1; ++it) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* a_dst = smem + stage * p.stage_bytes;
                uint8_t* b_dst = a_dst + WG_A_BYTES;
                mbar_expect_tx(&full[stage], tx);
                tma_load_2d(a_dst, &tmA, &full[stage], p0, it * WG_BK);
                tma_load_2d(a_dst + WG_BK * 128, &tmA, &full[stage], p0 + 64, it * WG_BK);
                for (int a = 0; a < p.q_atoms; ++a)
                    tma_load_2d(b_dst + a * WG_BK * 128, &tmB, &full[stage], q0 + a * 64, it * WG_BK);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // whole warp in the loop, one elected lane issues; descriptors advance by adds in the 16-byte address field
        const uint64_t dA0 = umma_desc_mnmajor_sw128(smem_u32(smem), WG_BK * 128);
        const uint64_t dB0 = umma_desc_mnmajor_sw128(smem_u32(smem) + WG_A_BYTES, WG_BK * 128);
        const uint32_t idesc = p.idesc, stage_step = static_cast<uint32_t>(p.stage_bytes) >> 4;
        int stage = 0;
        uint32_t phase = 0, off = 0, acc = 0;
        for (int it = it0; it < it1; ++it) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < WG_BK / 16; ++k)  // 16 tokens per MMA = two 8-token groups = 2048 B
                    umma_f16(tmem_base, dA0 + off + k * 128, dB0 + off + k * 128, idesc, (acc | k) ? 1u : 0u);
                umma_commit(&empty[stage]);
                if (it + 1 == it1) umma_commit(tfull);
            }
            __syncwarp();
            acc = 1;
            off += stage_step;
            if (++stage == p.stages) { stage = 0; phase ^= 1; off = 0; }
        }
    } else {
        const int lane_grp = warp & 3;
        const int r = lane_grp * 32 + lane;  // row of the P tile
        // slice layout [Q_pad / 4][P_pad rows][4 floats]: the 32 lanes (rows) of a warp write 512 contiguous bytes
        const long long P_pad = static_cast<long long>(p.p_tiles) * 128, Q_pad = static_cast<long long>(p.q_tiles) * p.q_tile;
        float* dst = p.ws + static_cast<long long>(split) * P_pad * Q_pad + (static_cast<long long>(q0 >> 2) * P_pad + p0 + r) * 4;
        if (p.direct) {
            mbar_wait(tfull, 0);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
            // every MMA has completed, so the operand ring is free: each epilogue warp transposes its 32 x 32 chunks through
            // a [32][36]-float patch of it (16-byte accesses, conflict free both ways) and touches `out` as 128-byte row
            // segments, four rows per instruction
            float* patch = reinterpret_cast<float*>(smem) + lane_grp * (32 * 36);
            const int sub_r = lane >> 3, c4 = lane & 7;
            const int row_base = p0 + lane_grp * 32;
            for (int c = 0; c < p.q_tile; c += 32) {
                if (q0 + c >= p.Q) break;
                uint32_t raw[32];
                tmem_ld_32x32(t_row + c, raw);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(patch + lane * 36 + 4 * q) =
                        make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                                    __uint_as_float(raw[4 * q + 3]));
                __syncwarp();
                const int col = q0 + c + 4 * c4;
                const bool col_ok = col < p.Q;  // Q % 8 == 0: a 4-vector is inside or outside
                float4 o[8];
                if (p.beta != 0.f) {
#pragma unroll
                    for (int it = 0; it < 8; ++it) {
                        const int row = row_base + 4 * it + sub_r;
                        o[it] = (col_ok && row < p.P) ? *reinterpret_cast<const float4*>(p.out + row * p.ldo + col)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int row = row_base + 4 * it + sub_r;
                    float4 v = *reinterpret_cast<const float4*>(patch + (4 * it + sub_r) * 36 + 4 * c4);
                    v.x *= p.alpha; v.y *= p.alpha; v.z *= p.alpha; v.w *= p.alpha;
                    if (p.beta != 0.f) { v.x += p.beta * o[it].x; v.y += p.beta * o[it].y; v.z += p.beta * o[it].z; v.w += p.beta * o[it].w; }
                    if (col_ok && row < p.P) *reinterpret_cast<float4*>(p.out + row * p.ldo + col) = v;
                }
                __syncwarp();
            }
        } else if (it1 > it0) {
            mbar_wait(tfull, 0);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16);
            for (int c = 0; c < p.q_tile; c += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(t_row + c, raw);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(dst + (static_cast<long long>((c >> 2) + q) * P_pad) * 4) =
                        make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                                    __uint_as_float(raw[4 * q + 3]));
            }
        } else {  // empty split (token count not divisible): contribute zeros
            for (int c = 0; c < p.q_tile; c += 4)
                *reinterpret_cast<float4*>(dst + (static_cast<long long>(c >> 2) * P_pad) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { __syncwarp(); tmem_dealloc(tmem_base, 256); }
}

// out[p, q] = alpha * sum_s ws[s][p][q] + beta * out[p, q]   (fixed summation order)
// block = 32 rows x 32 columns: the slices are read rows-fastest (512 contiguous bytes per warp, the layout the GEMM
// epilogue wrote), transposed through shared memory, and `out` is read/written as 128-byte row segments.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ out, int P, int Q, int P_pad, int Q_pad,
                    int splits, long long ldo, float alpha, float beta) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ float4 tile[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = blockIdx.x * 32, c4_0 = blockIdx.y * 8;
    const int q4s = Q >> 2;  // Q is a multiple of 8
    {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const int pr = r0 + tx, q4 = c4_0 + ty;
        if (pr < P && q4 < q4s) {
            const long long slice = static_cast<long long>(P_pad) * Q_pad;
            const float* src = ws + (static_cast<long long>(q4) * P_pad + pr) * 4;
            for (int s = 0; s < splits; ++s) {
                const float4 v = *reinterpret_cast<const float4*>(src + s * slice);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        tile[ty][tx] = acc;
    }
    __syncthreads();
    const int row = threadIdx.x >> 3, c4 = threadIdx.x & 7;
    const int pr = r0 + row, q4 = c4_0 + c4;
    if (pr >= P || q4 >= q4s) return;
    float4 acc = tile[c4][row];
    float* o = out + pr * ldo + 4 * q4;
    if (beta != 0.f) { acc.x = alpha * acc.x + beta * o[0]; acc.y = alpha * acc.y + beta * o[1]; acc.z = alpha * acc.z + beta * o[2]; acc.w = alpha * acc.w + beta * o[3]; }
    else { acc.x *= alpha; acc.y *= alpha; acc.z *= alpha; acc.w *= alpha; }
    o[0] = acc.x; o[1] = acc.y; o[2] = acc.z; o[3] = acc.w;
}

}  // namespace ctrl

using namespace ctrl;

extern "C" int ctrlora_wgrad_tn_f16(const void* a, long long lda, const void* b, long long ldb, int m, int p_dim, int q_dim,
                                    float* out, long long ldo, float alpha, float beta, float* ws, long long ws_bytes,
                                    void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!a || !b || !out || !ws || p_dim % 8 || q_dim % 8 || lda % 8 || ldb % 8 || m <= 0) return CTRLORA_ERR_ARG;
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.P = p_dim; p.Q = q_dim; p.M = m;
    p.q_tile = q_dim >= 256 ? 256 : ((q_dim + 63) / 64) * 64;
    p.q_atoms = p.q_tile / 64;
    p.p_tiles = (p_dim + 127) / 128;
    p.q_tiles = (q_dim + p.q_tile - 1) / p.q_tile;
    const int k_total = (m + WG_BK - 1) / WG_BK;
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    // enough token splits to fill the machine once, at least 8 k-iterations each, within the workspace
    int splits = (sms + p.p_tiles * p.q_tiles - 1) / (p.p_tiles * p.q_tiles);  // one wave: half the workspace traffic of two
    if (splits > k_total / 8) splits = k_total / 8;
    if (splits < 1) splits = 1;
    const long long slice = static_cast<long long>(p.p_tiles) * 128 * p.q_tiles * p.q_tile * 4;
    if (splits > 1) {
        if (slice > ws_bytes) splits = 1;  // does not fit the workspace: single split, written directly
        else if (splits * slice > ws_bytes) splits = static_cast<int>(ws_bytes / slice);
    }
    p.kiters_per_split = (k_total + splits - 1) / splits;
    p.splits = (k_total + p.kiters_per_split - 1) / p.kiters_per_split;
    p.direct = (p.splits == 1 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    p.out = out; p.ldo = ldo; p.alpha = alpha; p.beta = beta;
    if (!p.direct && p.splits * slice > ws_bytes) return CTRLORA_ERR_ARG;
    p.stage_bytes = WG_A_BYTES + p.q_atoms * WG_BK * 128;
    p.idesc = umma_idesc_f16(128, p.q_tile, 0) | (1u << 15) | (1u << 16);  // A and B MN-major
    p.ws = ws;
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[2] = {(uint64_t)p_dim, (uint64_t)m};
        uint64_t str[1] = {(uint64_t)lda * 2};
        uint32_t box[2] = {64, WG_BK};
        int rc = make_tmap_f16(&tmA, a, 2, dims, str, box);
        if (rc) return rc;
        uint64_t dimsb[2] = {(uint64_t)q_dim, (uint64_t)m};
        uint64_t strb[1] = {(uint64_t)ldb * 2};
        rc = make_tmap_f16(&tmB, b, 2, dimsb, strb, box);
        if (rc) return rc;
    }
    p.stages = (200 * 1024) / p.stage_bytes;
    if (p.stages > WG_STAGES) p.stages = WG_STAGES;
    const int smem_bytes = p.stages * p.stage_bytes + 1024 + 256;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(wgrad_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1280) != cudaSuccess)
            return CTRLORA_ERR_CUDA;
        attr = true;
    }
    if (launch_pdl(wgrad_tn_kernel, dim3(p.p_tiles * p.q_tiles, p.splits), dim3(WG_THREADS), (size_t)smem_bytes, stream, tmA,
                   tmB, p) != cudaSuccess)
        return CTRLORA_ERR_CUDA;
    if (p.direct) return cudaGetLastError() == cudaSuccess ? CTRLORA_OK : CTRLORA_ERR_CUDA;
    if (launch_pdl(wgrad_reduce_kernel, dim3((unsigned)((p_dim + 31) / 32), (unsigned)((q_dim / 4 + 7) / 8)), dim3(256), (size_t)0, stream,
                   (const float*)ws, out, p_dim, q_dim, p.p_tiles * 128, p.q_tiles * p.q_tile, p.splits, ldo, alpha, beta) !=
        cudaSuccess)
        return CTRLORA_ERR_CUDA;
    return cudaGetLastError() == cudaSuccess ? CTRLORA_OK : CTRLORA_ERR_CUDA;
}
