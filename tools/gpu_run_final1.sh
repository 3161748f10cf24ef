#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/r2f1_tests.log 2>&1
echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" gpurun_out/r2f1_tests.log | tail -8
python __graft_entry__.py 2>&1 | tail -2
python bench.py > gpurun_out/r2f1_bench_full.json 2> gpurun_out/r2f1_bench_full.err
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2f1_bench_$i.json 2> gpurun_out/r2f1_bench_$i.err; done
for f in gpurun_out/r2f1_bench_full.json gpurun_out/r2f1_bench_1.json gpurun_out/r2f1_bench_2.json; do grep "^{" $f | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); t=d['train']; a=d.get('roofline_attention',{}); print('sample', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'train', round(t['value'],1), round(t['ms_per_step'],2), 'train e2e', round(t['e2e']['value'],1), 'roof', round(d['roofline']['frac'],3), 'attn', round(a.get('frac',0),3), round(a.get('us_per_launch',0),1), 'cpu', d.get('cpu_baseline',{}).get('value'))"; done
python bench.py --workload pretrain --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2f1_pretrain.json
python -c "
import json; d=json.load(open('gpurun_out/r2f1_pretrain.json')); print('pretrain', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f1_launches_ddim_step.csv python tools/profile_step.py > gpurun_out/r2f1_ncu_step.log 2>&1
echo "ncu rc=$?"
python tools/profile_step.py --gemm-json gpurun_out/r2f1_gemm_shapes.json 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f1_launches_train_step.csv python tools/profile_train.py > gpurun_out/r2f1_ncu_train.log 2>&1
echo "ncu train rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:attention_stream64s -s 1 -c 1 -o gpurun_out/prof_attn_stream64s_r2 -f python tools/one_attn.py > gpurun_out/r2f1_ncu_attn.log 2>&1
echo "ncu attn rc=$?"
