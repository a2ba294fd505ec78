#!/bin/bash
# round 2, pass P: full ncu captures of the two ODE kernels at the icml size (after the SFU sigmoid / PDL changes)
B="python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads"
timeout 300 bash tools/gpu_ncu_cmd.sh r02_ws_bwd_icml elbo_bwd_ws 6 $B
timeout 300 bash tools/gpu_ncu_cmd.sh r02_team_fwd_icml elbo_fwd_team 6 $B
rm -f gpurun_out/*_details.csv
