#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_train_gpu.py tests/test_pretrain_gpu.py -q -s > gpurun_out/r2n_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2n_tests.log | tail -5; grep -n "rank [0-9]*: eps" gpurun_out/r2n_tests.log
python tools/debug_e2e.py > gpurun_out/r2n_e2e.log 2>&1; tail -24 gpurun_out/r2n_e2e.log
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('train', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))"; done
