#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2m_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2m_tests.log | tail -8; grep -n "rank [0-9]*: eps" gpurun_out/r2m_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
echo "bench rc=$?"; grep "^{" gpurun_out/r2m_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d['train']; print('sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2), 'traffic', d['roofline'].get('traffic'))"
tail -3 gpurun_out/r2m_bench.err
