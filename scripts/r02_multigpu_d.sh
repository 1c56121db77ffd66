#!/bin/bash
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/multigpu_check.py 2>&1 | grep -v "^\[W\|^W1\|\*\*\*\*\|OMP_NUM" | grep -E "it3|launch|ok|Error|error|assert|Traceback" | head -30
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3
