#!/bin/bash
# One GPU session: tests, bench line, reference arm, ncu launch list of the same command, one full ncu capture of the
# rollout kernel, clock64 phase stamps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 20000 --warmup 100 > gpurun_out/bench_native.json 2> gpurun_out/bench_native.err; cat gpurun_out/bench_native.json; tail -5 gpurun_out/bench_native.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.json 2>/dev/null; cat gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 150 --warmup 10 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 20 -c 2 -o gpurun_out/prof_rollout -f python bench.py --steps 40 --warmup 10 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python scripts/phase_stamps.py > gpurun_out/phase_stamps.txt 2>&1; tail -3 gpurun_out/phase_stamps.txt
