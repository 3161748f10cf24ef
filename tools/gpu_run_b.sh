#!/bin/bash
mkdir -p gpurun_out
python tools/debug_graph_stale.py > gpurun_out/r2b_debug.log 2>&1; tail -5 gpurun_out/r2b_debug.log
python -m pytest tests/test_vae_gpu.py tests/test_pretrain_gpu.py tests/test_train_gpu.py tests/test_variants_reference_gpu.py tests/test_model_gpu.py -q -s > gpurun_out/r2b_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED\|Error" gpurun_out/r2b_tests.log | tail -12
python bench.py --workload pretrain --steps 10 --warmup 3 > gpurun_out/r2b_pretrain.json 2> gpurun_out/r2b_pretrain.err
echo "pretrain bench rc=$?"; tail -c 1800 gpurun_out/r2b_pretrain.json; tail -5 gpurun_out/r2b_pretrain.err
