// Microbenchmark: cost of issuing / executing back-to-back tcgen05.mma (kind::f16, M=128, K=16) as a function of N,
// one issuing thread per CTA, 1 or 2 CTAs per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -I ../../ctrlora_b200/csrc
#include "common.cuh"
#include <cstdio>
using namespace ctrl;

__global__ void __launch_bounds__(128) k(int n, int reps, long long* out, int ab_same) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 64 * 1024);
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tptr, 256);
    for (int i = threadIdx.x; i < 16 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tptr;
    if (warp == 1) {
        const uint64_t da = umma_desc_kmajor_sw128(smem_u32(smem)), db = umma_desc_kmajor_sw128(smem_u32(smem) + 32768);
        const uint32_t idesc = umma_idesc_f16(128, n, 0);
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            t0 = clock64();
            if (ab_same) { for (int i = 0; i < reps; ++i) umma_f16_ts(tm, tm + 128 + 8 * (i & 3), db + 2 * (i & 3), idesc, 1u); }
            else { for (int i = 0; i < reps; ++i) umma_f16(tm, da + 2 * (i & 3), db + 2 * (i & 3), idesc, 1u); }
            t1 = clock64();
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        t2 = clock64();
        if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; }
        if (blockIdx.x == 0) { long long v = __shfl_sync(0xffffffffu, t0, 0); if (lane == 0) out[1] = t2 - (v ? v : t0); }
        // elected lane may not be lane 0: reduce max of (t1 - t0)
        long long d = t1 - t0;
        for (int o = 16; o; o >>= 1) { long long x = __shfl_xor_sync(0xffffffffu, d, o); d = x > d ? x : d; }
        long long e = t0; for (int o = 16; o; o >>= 1) { long long x = __shfl_xor_sync(0xffffffffu, e, o); e = x > e ? x : e; }
        if (lane == 0 && blockIdx.x == 0) { out[0] = d; out[1] = t2 - e; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 512;
    for (int ts = 0; ts < 2; ++ts)
    for (int ctas = 148; ctas <= 296; ctas += 148)
        for (int n : {16, 48, 64, 128}) {
            long long h[2];
            for (int it = 0; it < 2; ++it) {
                k<<<ctas, 128, 67 * 1024 + 1024>>>(n, reps, d, ts);
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            }
            cudaError_t e = cudaGetLastError();
            printf("%s ctas/SM=%d N=%3d: issue %6.1f cyc/MMA, issue+exec %6.1f cyc/MMA (floor N/2 = %d) %s\n", ts ? "TS" : "SS", ctas / 148, n,
                   (double)h[0] / reps, (double)h[1] / reps, n / 2 > 0 ? n / 2 : 0, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    return 0;
}
