#!/bin/bash
# round-2 GPU call A: parity suite (with printed error figures), baseline bench line, HBM-kernel timings + ncu counters
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -x > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; head -c 1500 gpurun_out/r2a_bench.json
python tools/profile_kernels.py norm ddim > gpurun_out/r2a_norm_times.log 2>&1
cat gpurun_out/r2a_norm_times.log
CTRLORA_PROFILE_ONCE=1 timeout 600 ncu --set full --clock-control none -k regex:"gn_|layernorm|ddim_update" -c 40 -f -o gpurun_out/r2a_hbm \
    python tools/profile_kernels.py norm ddim > gpurun_out/r2a_ncu.log 2>&1
echo "ncu rc=$?"
