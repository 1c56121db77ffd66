#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29515 scripts/exchange_stamps.py 300 2>&1 | grep -v "^\[W\|^W1\|\*\*\*\*\|OMP_NUM" | tee gpurun_out/exchange_stamps_n$N.txt | tail -20
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-2000} --warmup 20 > gpurun_out/bench_c2w_n${N}_p2p.json 2> gpurun_out/bench_c2w_n${N}_p2p.err
grep -E "Error|error" gpurun_out/bench_c2w_n${N}_p2p.err | head -3; python scripts/bench_summary.py < gpurun_out/bench_c2w_n${N}_p2p.json
BNV_DEBUG_DISABLE=8192 timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps ${STEPS:-2000} --warmup 20 2>/dev/null | python scripts/bench_summary.py
