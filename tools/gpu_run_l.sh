#!/bin/bash
# 8-GPU training scaling: all-reduce overlap variants
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { CTRLORA_ALLREDUCE_CUTS="$2" $TR --master-port $3 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/r2l_$1.json 2> gpurun_out/r2l_$1.err
  grep "^{" gpurun_out/r2l_$1.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$1', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1))" || tail -3 gpurun_out/r2l_$1.err; }
run none "" 29531
run ib3 "ib3" 29532
run all "middle,ib9,ib6,ib3" 29533
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('1gpu', round(d['value'],1), round(d['ms_per_step'],2))"
$TR --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 3 --workload pretrain > gpurun_out/r2l_pretrain8.json 2> gpurun_out/r2l_pretrain8.err
grep "^{" gpurun_out/r2l_pretrain8.json | head -c 330; tail -2 gpurun_out/r2l_pretrain8.err
