#!/bin/bash
mkdir -p gpurun_out
python tools/profile_kernels.py norm > gpurun_out/r2f_norm_times.log 2>&1; cat gpurun_out/r2f_norm_times.log
python tools/debug_determinism.py sd15 > gpurun_out/r2f_det_sd15.log 2>&1; tail -3 gpurun_out/r2f_det_sd15.log
python -m pytest tests -m gpu -q -s > gpurun_out/r2f_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2f_tests.log | tail -8
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; grep "^{" gpurun_out/r2f_bench.json | head -c 400; tail -3 gpurun_out/r2f_bench.err
CTRLORA_GN_CLUSTER=0 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload sample > gpurun_out/r2f_bench_nocluster.json 2>/dev/null
grep "^{" gpurun_out/r2f_bench_nocluster.json | head -c 300
