#!/bin/bash
# N GPUs: phase times of the exchange kernel inside real steps, then the N-rank checks, tests and bench
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29654 tools/peer_timing.py 2>&1 | grep "median us"
