#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -q -x -k "attention or attn" 2>&1 | tail -3
CTRLORA_ATTN_CROSS=0 timeout 300 python tools/time_attn_cross.py 2>&1 | grep "CROSS="
CTRLORA_ATTN_CROSS=1 timeout 300 python tools/time_attn_cross.py 2>&1 | grep "CROSS="
