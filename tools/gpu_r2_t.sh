#!/bin/bash
# 2 GPUs: N-rank exchange check (incl. NaN collective, sharded-vs-global), GPU tests, bench at N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/peer_check.py 2>&1 | grep -v Warning | tail -12
bash tools/gpu_r2_i.sh
