#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py tests/test_style_ip_gpu.py -q > gpurun_out/r2c3_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2c3_tests.log | tail -8
CTRLORA_ATTN_CROSS=1 timeout 300 python tools/time_attn_cross.py 2>&1 | grep "CROSS="
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2c3_bench.json
python -c "
import json
d=json.load(open('gpurun_out/r2c3_bench.json')); t=d['train']; print('sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2), 'roof', d['roofline']['frac'])"
