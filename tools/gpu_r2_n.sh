#!/bin/bash
# round 2, pass N: full ncu captures of the three small launches around the ODE kernels at the icml size
B="python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads"
timeout 300 bash tools/gpu_ncu_cmd.sh r02_enc_fwd_icml enc_fwd_kernel 6 $B
timeout 300 bash tools/gpu_ncu_cmd.sh r02_enc_bwd_icml enc_bwd_kernel 6 $B
timeout 300 bash tools/gpu_ncu_cmd.sh r02_enc_adam_icml enc_lin_wgrad_adam 6 $B
rm -f gpurun_out/*_details.csv
