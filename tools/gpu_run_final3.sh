#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2f3_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2f3_tests.log | tail -6
python __graft_entry__.py 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f3_bench.json 2> gpurun_out/r2f3_bench.err
grep "^{" gpurun_out/r2f3_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d['train']; a=d.get('roofline_attention',{}); print('sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2), 'roof', round(d['roofline']['frac'],3), 'attn', round(a.get('frac',0),3), 'spread', t.get('replica_param_spread'))"
python bench.py --workload pretrain --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('pretrain', round(d['value'],1), round(d['ms_per_step'],2))"
