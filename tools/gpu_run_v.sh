#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_train_kernels_gpu.py -q -x -k "attention or attn" > gpurun_out/r2v_tests.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2v_tests.log
CTRLORA_ATTN_STREAM64=0 timeout 300 python tools/time_attn.py 2>&1 | grep STREAM64
CTRLORA_ATTN_STREAM64=1 timeout 300 python tools/time_attn.py 2>&1 | grep STREAM64
