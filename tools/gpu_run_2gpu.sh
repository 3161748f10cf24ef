#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_2gpu_bench.json 2> gpurun_out/r2_2gpu_bench.err
echo "rc=$?"; grep "^{" gpurun_out/r2_2gpu_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d['train']; print('2gpu sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2), 'n', d['n_gpus'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload pretrain --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_2gpu_pretrain.err | grep "^{" > gpurun_out/r2_2gpu_pretrain.json
python -c "
import json; d=json.load(open('gpurun_out/r2_2gpu_pretrain.json')); print('2gpu pretrain', round(d['value'],1), round(d['ms_per_step'],2))"
tail -3 gpurun_out/r2_2gpu_bench.err
