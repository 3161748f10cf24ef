#!/bin/bash
for v in 0 1 0 1; do
CTRLORA_ATTN_CROSS=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d['train']; print('CROSS=$v sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2), 'roof', round(d['roofline']['frac'],4))"
done
