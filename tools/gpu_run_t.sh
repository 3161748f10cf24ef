#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_round2_kernels_gpu.py tests/test_kernels_gpu.py tests/test_train_gpu.py tests/test_pretrain_gpu.py -q > gpurun_out/r2t_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2t_tests.log | tail -12
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d['train']; print('sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2))"
python bench.py --workload pretrain --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2t_pretrain.json
python -c "
import json; d=json.load(open('gpurun_out/r2t_pretrain.json')); print('pretrain', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['frac'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2t_launches_pretrain_step.csv python tools/profile_pretrain.py > gpurun_out/r2t_ncu_pretrain.log 2>&1
echo "ncu pretrain rc=$?"
