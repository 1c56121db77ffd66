#!/bin/bash
# GPU session: parity suite, bench (native + reference arm), ncu launch list of the bench command, full ncu captures
# (main rollout kernel; batched and stochastic variants; Monte-Carlo risk map), phase stamps, widened-path rates.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
python bench.py --steps 20000 --warmup 100 > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err; cat gpurun_out/bench_native.json; tail -3 gpurun_out/bench_native.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 150 --warmup 10 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 20 -c 2 -o gpurun_out/prof_rollout -f python bench.py --steps 40 --warmup 10 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 6 -c 2 -o gpurun_out/prof_batch -f python scripts/profile_targets.py batch > gpurun_out/ncu_batch.log 2>&1; tail -1 gpurun_out/ncu_batch.log
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 6 -c 2 -o gpurun_out/prof_stoch -f python scripts/profile_targets.py stoch > gpurun_out/ncu_stoch.log 2>&1; tail -1 gpurun_out/ncu_stoch.log
ncu --set full --clock-control none --import-source on -k regex:risk_mc_kernel -s 2 -c 1 -o gpurun_out/prof_risk -f python scripts/profile_targets.py risk 4 > gpurun_out/ncu_risk.log 2>&1; tail -1 gpurun_out/ncu_risk.log
python scripts/phase_stamps.py > gpurun_out/phase_stamps.txt 2>&1; tail -3 gpurun_out/phase_stamps.txt
timeout 600 python scripts/bench_configs.py > gpurun_out/bench_configs.jsonl 2>&1; tail -4 gpurun_out/bench_configs.jsonl
ls -la gpurun_out
