#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2g_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2g_tests.log | tail -8
python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; grep "^{" gpurun_out/r2g_bench.json | head -c 300; tail -3 gpurun_out/r2g_bench.err
python bench.py --impl reference-gpu --steps 5 --warmup 3 > gpurun_out/r2g_refgpu.json 2> gpurun_out/r2g_refgpu.err
echo "refgpu rc=$?"; grep "^{" gpurun_out/r2g_refgpu.json | head -c 900; tail -3 gpurun_out/r2g_refgpu.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2g_launches_ddim_step.csv python tools/profile_step.py > gpurun_out/r2g_ncu_step.log 2>&1
echo "ncu rc=$?"
