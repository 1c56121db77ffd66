#!/bin/bash
mkdir -p gpurun_out
BNV_DEBUG_DISABLE=4096 python scripts/phase_stamps.py 2>&1 | tail -3
BNV_DEBUG_DISABLE=4096 python scripts/phase_stamps.py 131072 2>&1 | tail -2
BNV_DEBUG_DISABLE=4096 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -o gpurun_out/prof_wide -f python scripts/phase_stamps.py > gpurun_out/ncu_wide.log 2>&1; tail -2 gpurun_out/ncu_wide.log
