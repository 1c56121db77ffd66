#!/bin/bash
# two-GPU session: sharded-solver parity (NCCL + fused P2P exchange) and the N=2 bench lines for both exchange modes
timeout 300 python -m pytest tests/test_multigpu_gpu.py -q -x 2>&1 | tail -3
for ex in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3000 --warmup 50 --exchange $ex 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$ex', 'value',round(d['value']), 'us/step',round(d['ms_per_step']*1e3,2), d['config']['parallelism'])"
done
