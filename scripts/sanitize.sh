#!/bin/bash
# compute-sanitizer passes over small parity cases of every kernel mode (memcheck, racecheck, synccheck)
mkdir -p gpurun_out
SEL='prelaunched or golden_single or stochastic_golden or batched_matches or env_step or collision or risk_map_golden or dwa or without_a_staged or graph or top_samples'
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_parity_gpu.py tests/test_ext_gpu.py -m gpu -q -x -k "$SEL" 2>&1 | tail -6 | tee gpurun_out/sanitize_$tool.txt
done
