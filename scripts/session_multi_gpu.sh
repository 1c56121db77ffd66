#!/bin/bash
# N-GPU session: parity of both exchanges + host-driven path, exchange stamps, bench lines (driver-style short run and
# a long run of the weak-scaling line; nccl exchange; strong c2; environment-sharded c3), reference arm
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/multigpu_check.py 2>&1 | grep -v "^\[W\|^W1\|\*\*\*\*\|OMP_NUM" | grep -E "it3|launch|ok|Error|error|assert|Traceback" | head -30 | tee gpurun_out/multigpu_check_n$N.txt
timeout 600 $TR --master-port 29515 scripts/exchange_stamps.py 300 2>&1 | grep -v "^\[W\|^W1\|\*\*\*\*\|OMP_NUM\|NCCL version" | tee gpurun_out/exchange_stamps_n$N.txt | tail -16
timeout 600 $TR --master-port 29516 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c2w_n${N}_driverstyle.json 2> gpurun_out/bench_c2w_n${N}_driverstyle.err
echo -n "driver-style: "; python scripts/bench_summary.py < gpurun_out/bench_c2w_n${N}_driverstyle.json
for ex in p2p nccl; do
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps ${STEPS:-2000} --warmup 20 --exchange $ex > gpurun_out/bench_c2w_n${N}_$ex.json 2> gpurun_out/bench_c2w_n${N}_$ex.err
  grep -E "Error|error" gpurun_out/bench_c2w_n${N}_$ex.err | head -3; echo -n "$ex: "; python scripts/bench_summary.py < gpurun_out/bench_c2w_n${N}_$ex.json
done
for c in c2 c3; do
  timeout 600 $TR --master-port 29513 bench.py --gpus $N --config $c --steps ${STEPS:-2000} --warmup 20 > gpurun_out/bench_${c}_n${N}.json 2> gpurun_out/bench_${c}_n${N}.err
  grep -E "Error|error" gpurun_out/bench_${c}_n${N}.err | head -3; python scripts/bench_summary.py < gpurun_out/bench_${c}_n${N}.json
done
timeout 300 $TR --master-port 29514 bench.py --gpus $N --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_c2w_n${N}_reference.json; cut -c1-300 gpurun_out/bench_c2w_n${N}_reference.json
