#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_stream64 -s 1 -c 1 -o gpurun_out/prof_attn_stream64_r2w -f python tools/one_attn.py > gpurun_out/r2w_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2w_ncu.log
