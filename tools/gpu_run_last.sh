#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:attention_cross -s 1 -c 1 -o gpurun_out/prof_attn_cross_r2 -f python tools/one_attn_cross.py > gpurun_out/r2last_ncu.log 2>&1
echo "ncu rc=$?"
timeout 200 python bench.py --impl reference-gpu --steps 3 --warmup 1 2>/dev/null | grep "^{" | cut -c1-400
