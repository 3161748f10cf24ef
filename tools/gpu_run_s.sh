#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_round2_kernels_gpu.py tests/test_kernels_gpu.py tests/test_train_gpu.py tests/test_pretrain_gpu.py -q > gpurun_out/r2s_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2s_tests.log | tail -5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2s_launches_pretrain_step.csv python tools/profile_pretrain.py > gpurun_out/r2s_ncu_pretrain.log 2>&1
echo "ncu pretrain rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d['train']; print('sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2))"
