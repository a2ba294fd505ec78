#!/bin/bash
# round 2, pass G: lane-split forward kernel: parity with the default kernels, kernel-only timing, ncu of both
timeout 600 python tools/lane_check.py 2>&1 | tail -7
for lane in 0 1; do VIHDS_FWD_LANE=$lane timeout 300 python tools/microbench.py --B 36 --IW 200 --T 86 2>&1 | tail -2; done
for lane in 0 1; do VIHDS_FWD_LANE=$lane timeout 300 python tools/microbench.py --B 1024 --IW 128 --T 500 --iters 3 2>&1 | tail -2; done
export VIHDS_FWD_LANE=1
bash tools/gpu_ncu_cmd.sh r02_lane_fwd_icml elbo_fwd_lane 1 python tools/microbench.py --B 36 --IW 200 --T 86 --iters 1 > /dev/null
