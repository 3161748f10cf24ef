#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2o_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2o_tests.log | tail -8
for i in 1 2 3; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2o_bench_$i.json 2> gpurun_out/r2o_bench_$i.err
grep "^{" gpurun_out/r2o_bench_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d['train']; print('sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2), 'train e2e', round(t['e2e']['value'],1))"; done
python __graft_entry__.py 2>&1 | tail -2
