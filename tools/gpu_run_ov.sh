#!/bin/bash
mkdir -p gpurun_out
run() {
CTRLORA_PRETRAIN_ALLREDUCE_CUTS="$1" CTRLORA_OVERLAP_SM_RESERVE="$2" NCCL_MAX_NCHANNELS="$3" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --workload pretrain --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2ov_pretrain.err | grep "^{" > gpurun_out/r2ov_pretrain.json
python -c "
import json; d=json.load(open('gpurun_out/r2ov_pretrain.json')); print('cuts=[$1] reserve=$2 nch=$3: 2gpu pretrain', round(d['value'],1), round(d['ms_per_step'],2))"
}
run "" 16 ""
run "ib9,ib6" 0 ""
run "ib9,ib6" 32 ""
run "ib9,ib6" 8 8
run "middle,ib9,ib6,ib3" 16 ""
run "ib9" 16 ""
