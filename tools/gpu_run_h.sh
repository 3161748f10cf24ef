#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_train_kernels_gpu.py tests/test_variants_reference_gpu.py -q > gpurun_out/r2h_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2h_tests.log | tail -6
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload sample > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; grep "^{" gpurun_out/r2h_bench.json | head -c 260
CTRLORA_PROFILE_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -c 2 -f -o gpurun_out/r2h_gemm1x1 python tools/profile_kernels.py gemm > gpurun_out/r2h_ncu_gemm.log 2>&1
echo "ncu rc=$?"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2h_cpuref.json 2> gpurun_out/r2h_cpuref.err
echo "cpuref rc=$?"; grep "^{" gpurun_out/r2h_cpuref.json | head -c 1500; tail -4 gpurun_out/r2h_cpuref.err
