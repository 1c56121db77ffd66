N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/multigpu_check.py 2>&1 | grep -v "^\[W\|^W1\|\*\*\*\*" | grep -E "it[0-9]|ok|Error|error|assert" | head -20
for ex in p2p nccl; do timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3000 --warmup 50 --exchange $ex 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$ex value',round(d['value']), 'ms/step',round(d['ms_per_step']*1e3,2),'us kernel_us',round(d['roofline']['kernel_us'],2),'b2b',round(d['config']['back_to_back_ms_per_step']*1e3,2), d['config']['parallelism'], 'launches', d['gpu_launches'])"; done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-300
