#!/bin/bash
mkdir -p gpurun_out
echo "== default"; python tools/profile_kernels.py gemm 2>&1 | grep "^gemm" | tee gpurun_out/r2i_gemm_default.log
echo "== nbuf1"; CTRLORA_GEMM_EPI_NBUF=1 python tools/profile_kernels.py gemm 2>&1 | grep "^gemm" | tee gpurun_out/r2i_gemm_nbuf1.log
echo "== pair everywhere"; CTRLORA_GEMM_PAIR=2 python tools/profile_kernels.py gemm 2>&1 | grep "^gemm" | tee gpurun_out/r2i_gemm_pair2.log
CTRLORA_GEMM_EPI_NBUF=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload sample 2>/dev/null | grep "^{" | head -c 200
