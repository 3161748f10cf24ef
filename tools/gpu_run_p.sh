#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_variants_reference_gpu.py tests/test_variants_gpu.py tests/test_checkpoint_flow_gpu.py -q -s > gpurun_out/r2p_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2p_tests.log | tail -8; grep -n "inference apply_model" gpurun_out/r2p_tests.log
