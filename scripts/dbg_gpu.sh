timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python scripts/phase_stamps.py | tail -4
python bench.py --steps 5000 --warmup 50
