#!/bin/bash
mkdir -p gpurun_out
python tools/profile_kernels.py norm > gpurun_out/r2e_norm_times.log 2>&1; cat gpurun_out/r2e_norm_times.log
python tools/debug_determinism.py sd15 > gpurun_out/r2e_det_sd15.log 2>&1; tail -6 gpurun_out/r2e_det_sd15.log
python -m pytest tests -m gpu -q -s > gpurun_out/r2e_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2e_tests.log | tail -12
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; grep "^{" gpurun_out/r2e_bench.json | head -c 700; tail -3 gpurun_out/r2e_bench.err
