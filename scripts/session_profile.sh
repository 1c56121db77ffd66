#!/bin/bash
# Round-2 profiling session (one GPU): ncu launch lists of the bench command (c1, c2), full captures of the latency and
# the wide variant of the rollout kernel (+ batched / stochastic), summarised ON the box (gpurun_out/ is capped at
# 64 MiB: only the text summaries and the two main reports travel back); memcheck.
mkdir -p gpurun_out/summaries
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 150 --warmup 10 > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 100 --csv --log-file gpurun_out/launches_c2.csv python bench.py --config c2 --steps 30 --warmup 5 > gpurun_out/ncu_bench_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 20 -c 2 -o gpurun_out/prof_rollout -f python bench.py --steps 40 --warmup 10 > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 6 -c 1 -o gpurun_out/prof_wide -f python bench.py --config c2 --steps 10 --warmup 3 > gpurun_out/ncu_wide.log 2>&1; tail -1 gpurun_out/ncu_wide.log
ncu --set full --clock-control none -k regex:rollout_kernel -s 6 -c 1 -o gpurun_out/prof_batch -f python scripts/profile_targets.py batch > gpurun_out/ncu_batch.log 2>&1; tail -1 gpurun_out/ncu_batch.log
ncu --set full --clock-control none -k regex:rollout_kernel -s 6 -c 1 -o gpurun_out/prof_stoch -f python scripts/profile_targets.py stoch > gpurun_out/ncu_stoch.log 2>&1; tail -1 gpurun_out/ncu_stoch.log
python scripts/ncu_summary.py --tag r02 --out-dir gpurun_out/summaries > gpurun_out/ncu_summary.log 2>&1; tail -3 gpurun_out/ncu_summary.log
ncu -i gpurun_out/prof_wide.ncu-rep --page source --csv > gpurun_out/summaries/r02_wide_source_page.csv 2>/dev/null
ncu -i gpurun_out/prof_rollout.ncu-rep --page source --csv > gpurun_out/summaries/r02_rollout_source_page.csv 2>/dev/null
rm -f gpurun_out/prof_batch.ncu-rep gpurun_out/prof_stoch.ncu-rep gpurun_out/prof_rollout.ncu-rep
du -sh gpurun_out
