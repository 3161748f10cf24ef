#!/bin/bash
mkdir -p gpurun_out
for P in 0 8 9 6; do
CTRLORA_ATTN_POLY=$P timeout 300 python tools/time_attn.py 2>&1 | grep "STREAM64" | head -2 | sed "s/^/POLY=$P /"
done
CTRLORA_ATTN_POLY=6 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -q -x -k "attention or attn" 2>&1 | tail -3
CTRLORA_ATTN_POLY=8 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -q -x -k "attention or attn" 2>&1 | tail -3
