#!/bin/bash
mkdir -p gpurun_out
python tools/bench_extras.py > gpurun_out/r2q_extras.json 2> gpurun_out/r2q_extras.err
echo "rc=$?"; cat gpurun_out/r2q_extras.json | tail -40; tail -5 gpurun_out/r2q_extras.err
