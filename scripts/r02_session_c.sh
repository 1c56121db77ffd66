#!/bin/bash
# wide-variant iteration: forced-wide parity, stamps, bench of the large configurations
mkdir -p gpurun_out
BNV_DEBUG_DISABLE=4096 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_forced_wide.txt
BNV_DEBUG_DISABLE=4096 python scripts/phase_stamps.py 2>&1 | tail -1
python scripts/phase_stamps.py 131072 2>&1 | tail -1
for c in ${CONFIGS:-c2 c3 c4}; do
  timeout 600 python bench.py --config $c --steps ${STEPS:-500} --warmup 20 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -2 gpurun_out/bench_$c.err; python scripts/bench_summary.py < gpurun_out/bench_$c.json
done
BNV_DEBUG_DISABLE=4096 timeout 600 python bench.py --config c1 --steps 500 --warmup 20 2>/dev/null | python scripts/bench_summary.py
