#!/bin/bash
mkdir -p gpurun_out
CTRLORA_GN_CLUSTER=0 python tools/debug_determinism.py tiny > gpurun_out/r2c_det_tiny_2pass.log 2>&1; tail -14 gpurun_out/r2c_det_tiny_2pass.log
python tools/debug_determinism.py tiny > gpurun_out/r2c_det_tiny.log 2>&1; tail -14 gpurun_out/r2c_det_tiny.log
python tools/debug_determinism.py sd15 > gpurun_out/r2c_det_sd15.log 2>&1; tail -14 gpurun_out/r2c_det_sd15.log
python -m pytest tests -m gpu -q -s > gpurun_out/r2c_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2c_tests.log | tail -12
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?"; grep "^{" gpurun_out/r2c_bench.json | head -c 900
python tools/profile_kernels.py norm > gpurun_out/r2c_norm_times.log 2>&1; cat gpurun_out/r2c_norm_times.log
