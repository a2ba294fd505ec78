#!/bin/bash
# 8 GPUs: exchange phase times (low-latency form and flag form), then the scaling bench in both forms
bash tools/gpu_r2_u.sh 2>&1 | head -3
VIHDS_PEER_LL=0 bash tools/gpu_r2_u.sh 2>&1 | head -3
N=$(nvidia-smi -L | wc -l)
VIHDS_PEER_LL=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29658 bench.py --gpus $N --steps 100 --warmup 5 --no-extra-workloads --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('flags form:', d['ms_per_step'], d['value'], d['params_identical_across_ranks'])"
bash tools/gpu_r2_scale.sh
