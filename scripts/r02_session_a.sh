#!/bin/bash
# Round-2 single-GPU session: parity suite, bench lines of every configuration, phase stamps.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
for c in c1 c2 c3 c4; do
  timeout 600 python bench.py --config $c --steps ${STEPS:-1000} --warmup 20 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -2 gpurun_out/bench_$c.err; python scripts/bench_summary.py < gpurun_out/bench_$c.json
done
python scripts/phase_stamps.py > gpurun_out/phase_stamps.txt 2>&1; tail -2 gpurun_out/phase_stamps.txt
