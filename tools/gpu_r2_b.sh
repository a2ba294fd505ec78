#!/bin/bash
# round 2, pass B: ncu of the dr_blackbox tensor-core kernels at the icml size
export VIHDS_BB_IMPL=mma
bash tools/gpu_ncu_cmd.sh r02_bbm_fwd_icml bbm_fwd 1 python tools/bb_microbench.py --B 36 --IW 200 --iters 1
bash tools/gpu_ncu_cmd.sh r02_bbm_bwd_icml bbm_bwd 1 python tools/bb_microbench.py --B 36 --IW 200 --iters 1
