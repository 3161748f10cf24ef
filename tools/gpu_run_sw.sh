#!/bin/bash
mkdir -p gpurun_out
SWEEP_SMALL=1 timeout 600 python tools/sweep_gemm.py > gpurun_out/r2_sweep_gemm_small.log 2>&1
cat gpurun_out/r2_sweep_gemm_small.log | tail -60
