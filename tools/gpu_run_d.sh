#!/bin/bash
mkdir -p gpurun_out
python tools/profile_kernels.py norm > gpurun_out/r2d_norm_times.log 2>&1; cat gpurun_out/r2d_norm_times.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest "tests/test_pretrain_gpu.py::test_pretrain_loss_and_all_gradients_vs_reference_autograd" -x -q > gpurun_out/r2d_sanitizer.log 2>&1
echo "sanitizer rc=$?"; grep -n "Invalid\|at 0x\|by thread\|Error\|ERROR SUMMARY\|passed\|failed" gpurun_out/r2d_sanitizer.log | head -40
CTRLORA_GN_CLUSTER=0 python tools/debug_determinism.py tiny > gpurun_out/r2d_det_tiny_2pass.log 2>&1; tail -8 gpurun_out/r2d_det_tiny_2pass.log
python tools/debug_determinism.py tiny > gpurun_out/r2d_det_tiny.log 2>&1; tail -8 gpurun_out/r2d_det_tiny.log
python tools/debug_determinism.py sd15 > gpurun_out/r2d_det_sd15.log 2>&1; tail -8 gpurun_out/r2d_det_sd15.log
