#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2r_launches_ddim_step.csv python tools/profile_step.py > gpurun_out/r2r_ncu_step.log 2>&1
echo "ncu rc=$?"
python tools/profile_step.py --gemm-json gpurun_out/r2r_gemm_shapes.json 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2r_launches_train_step.csv python tools/profile_train.py > gpurun_out/r2r_ncu_train.log 2>&1
echo "ncu train rc=$?"
