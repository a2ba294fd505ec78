#!/bin/bash
# round 2, pass A: dr_blackbox tensor-core kernels vs the scalar kernels, then kernel-only timings of both
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python tools/bb_mma_check.py 2>&1 | tail -40
for impl in scalar mma; do
  VIHDS_BB_IMPL=$impl timeout 300 python tools/bb_microbench.py --B 36 --IW 200 2>&1 | tail -1
done
VIHDS_BB_IMPL=mma timeout 300 python tools/bb_microbench.py --B 1024 --IW 128 --iters 3 2>&1 | tail -1
VIHDS_BB_IMPL=scalar timeout 300 python tools/bb_microbench.py --B 1024 --IW 128 --iters 3 2>&1 | tail -1
