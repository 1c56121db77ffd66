#!/bin/bash
# Single-GPU session: parity suite (default kernel selection, then every eligible solver forced onto the
# wide variant), bench lines of every configuration, driver-style short run, reference arm, phase stamps, lean mode,
# closed loop.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.txt
BNV_DEBUG_DISABLE=4096 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_forced_wide.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.txt
for c in c1 c2 c3 c4; do
  timeout 600 python bench.py --config $c --steps 2000 --warmup 20 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err
  tail -2 gpurun_out/bench_$c.err; python scripts/bench_summary.py < gpurun_out/bench_$c.json
done
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c1_driverstyle.json 2>/dev/null; echo -n "driver-style: "; python scripts/bench_summary.py < gpurun_out/bench_c1_driverstyle.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_c1_reference.json 2>/dev/null; cut -c1-200 gpurun_out/bench_c1_reference.json
python scripts/phase_stamps.py 2>&1 | tail -2 | tee gpurun_out/phase_stamps.txt
python scripts/phase_stamps.py 131072 2>&1 | tail -1 | tee -a gpurun_out/phase_stamps.txt
python scripts/phase_stamps.py 32768 stoch 2>&1 | tail -1 | tee -a gpurun_out/phase_stamps.txt
python scripts/bench_lean.py 2>&1 | tee gpurun_out/bench_lean.jsonl
for m in "" "--unfused" "--no-graph" "--no-graph --unfused"; do python examples/closed_loop.py $m 2>&1 | head -1; done | tee gpurun_out/closed_loop.txt
python scripts/launch_floor.py 2>&1 | tee gpurun_out/launch_floor.txt
