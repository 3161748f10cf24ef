#!/bin/bash
mkdir -p gpurun_out
E2E_PROFILE=1 timeout 600 python tools/debug_e2e.py > gpurun_out/r2_e2e_profile.log 2>&1
grep -n "median" gpurun_out/r2_e2e_profile.log
