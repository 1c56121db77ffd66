#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for rep in 1 2; do
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-2000} --warmup 20 > gpurun_out/bench_c2w_n${N}_p2p.json 2> gpurun_out/bench_c2w_n${N}_p2p.err
grep -E "Error|error" gpurun_out/bench_c2w_n${N}_p2p.err | head -3; python scripts/bench_summary.py < gpurun_out/bench_c2w_n${N}_p2p.json
done
timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c2w_n${N}_driverstyle.json 2>/dev/null; echo -n "driver-style: "; python scripts/bench_summary.py < gpurun_out/bench_c2w_n${N}_driverstyle.json
timeout 600 $TR --master-port 29514 bench.py --gpus $N --config c2 --steps ${STEPS:-2000} --warmup 20 > gpurun_out/bench_c2_n${N}.json 2>/dev/null; python scripts/bench_summary.py < gpurun_out/bench_c2_n${N}.json
