#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_2gpu.json 2> gpurun_out/r2j_bench_2gpu.err
echo "bench2 rc=$?"; grep "^{" gpurun_out/r2j_bench_2gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('sample', d['value'], d['ms_per_step'], 'train', d['train']['value'], d['train']['ms_per_step'], d.get('train_images_per_sec'))"
tail -5 gpurun_out/r2j_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload pretrain --steps 10 --warmup 3 > gpurun_out/r2j_pretrain_2gpu.json 2> gpurun_out/r2j_pretrain_2gpu.err
echo "pretrain2 rc=$?"; grep "^{" gpurun_out/r2j_pretrain_2gpu.json | head -c 400; tail -5 gpurun_out/r2j_pretrain_2gpu.err
