#!/bin/bash
# round 2, pass C: blackbox check + timings, then the GPU test-suite
mkdir -p gpurun_out
timeout 600 python tools/bb_mma_check.py 2>&1 | tail -14
for impl in scalar mma; do
  VIHDS_BB_IMPL=$impl timeout 300 python tools/bb_microbench.py --B 36 --IW 200 2>&1 | tail -1
done
VIHDS_BB_IMPL=mma timeout 300 python tools/bb_microbench.py --B 1024 --IW 128 --iters 3 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
