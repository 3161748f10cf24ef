#!/bin/bash
mkdir -p gpurun_out
show() { grep "^{" $1 | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d.get('train',{})
print('$2', 'sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3) if 'ms_per_step' in d['e2e'] else '', 'train', round(t.get('value',0),1), round(t.get('ms_per_step',0),2))"; }
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload sample > gpurun_out/r2k_1gpu_s10.json 2>/dev/null; show gpurun_out/r2k_1gpu_s10.json 1gpu_s10
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_1gpu.json 2>/dev/null; show gpurun_out/r2k_1gpu.json 1gpu_s20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2k_2gpu.json 2>gpurun_out/r2k_2gpu.err; show gpurun_out/r2k_2gpu.json 2gpu_overlap
CTRLORA_OVERLAP_ALLREDUCE=0 $TR --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/r2k_2gpu_nooverlap.json 2>/dev/null
grep "^{" gpurun_out/r2k_2gpu_nooverlap.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('2gpu_train_nooverlap', round(d['value'],1), round(d['ms_per_step'],2))"
$TR --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/r2k_2gpu_train.json 2>/dev/null
grep "^{" gpurun_out/r2k_2gpu_train.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('2gpu_train_overlap', round(d['value'],1), round(d['ms_per_step'],2))"
CUDA_VISIBLE_DEVICES=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload train > gpurun_out/r2k_1gpu_train.json 2>/dev/null
grep "^{" gpurun_out/r2k_1gpu_train.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('1gpu_train', round(d['value'],1), round(d['ms_per_step'],2))"
