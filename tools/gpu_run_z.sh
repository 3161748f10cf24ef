#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_style_ip_gpu.py tests/test_variants_reference_gpu.py tests/test_model_gpu.py -q -s > gpurun_out/r2z_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED\|Error\|style\|IPCross" gpurun_out/r2z_tests.log | tail -20
