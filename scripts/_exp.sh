timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python scripts/phase_stamps.py 32768 2>&1 | tail -3 | cut -c1-200
python scripts/phase_stamps.py 32768 stoch 2>&1 | tail -3 | cut -c1-200
python scripts/bench_lean.py 2>&1 | sed -n 2,2p
timeout 600 python bench.py --config c4 --steps 2000 --warmup 20 2>/dev/null | python scripts/bench_summary.py
