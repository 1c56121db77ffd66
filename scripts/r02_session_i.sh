#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.txt
BNV_DEBUG_DISABLE=4096 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_forced_wide.txt
python scripts/phase_stamps.py 32768 stoch 2>&1 | tail -2
python scripts/phase_stamps.py 32768 2>&1 | tail -1
for c in ${CONFIGS:-c4 c1}; do
  timeout 600 python bench.py --config $c --steps ${STEPS:-2000} --warmup 20 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -2 gpurun_out/bench_$c.err; python scripts/bench_summary.py < gpurun_out/bench_$c.json
done
