#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for rep in 1 2; do
  timeout 600 python bench.py --config c4 --steps 1500 --warmup 30 2>/dev/null | python scripts/bench_summary.py | cut -c1-120
done
python scripts/bench_bigk.py 2>&1 | cut -c1-150
python scripts/phase_stamps.py 32768 stoch 2>&1 | tail -1
python scripts/phase_stamps.py 32768 2>&1 | tail -1
