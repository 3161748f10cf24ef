#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --workload pretrain --steps 12 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2ov2_pretrain.err | grep "^{" > gpurun_out/r2ov2_pretrain.json
python -c "
import json; d=json.load(open('gpurun_out/r2ov2_pretrain.json')); p=d['pretrain']; print('2gpu pretrain', round(d['value'],1), round(d['ms_per_step'],2), 'spread', p.get('replica_param_spread'), 'cuts', p.get('allreduce_cuts'), 'skipped', p.get('skipped_steps'))"
tail -3 gpurun_out/r2ov2_pretrain.err | cut -c1-300
