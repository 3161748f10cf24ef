#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2f2_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2f2_tests.log | tail -8
python __graft_entry__.py 2>&1 | tail -1
python bench.py > gpurun_out/r2f2_bench_full.json 2> gpurun_out/r2f2_bench_full.err
grep "^{" gpurun_out/r2f2_bench_full.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d['train']; a=d.get('roofline_attention',{}); print('sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2), 'train e2e', round(t['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'attn', round(a.get('frac',0),3), 'cpu', d.get('cpu_baseline',{}).get('value'), 'launches', d['gpu_launches'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f2_launches_ddim_step.csv python tools/profile_step.py > gpurun_out/r2f2_ncu_step.log 2>&1
echo "ncu rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f2_launches_train_step.csv python tools/profile_train.py > gpurun_out/r2f2_ncu_train.log 2>&1
echo "ncu train rc=$?"
