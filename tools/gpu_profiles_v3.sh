#!/bin/bash
# One GPU call that regenerates the r01_v3 evidence: bench lines of the four workloads, the launch list of bench.py,
# full ncu captures (csv-exported on the box) of the ODE kernels at the bench workload and at N = 131,072 x T = 500.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python bench.py 2>gpurun_out/v3_bench.err | tail -1 > gpurun_out/v3_bench_dr_constant_icml.json
for wl in synthetic_dr_constant dr_blackbox_icml relay_constant_precisions; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline 2>>gpurun_out/v3_bench.err | tail -1 > gpurun_out/v3_bench_$wl.json
done
timeout 120 python tools/microbench.py --B 36 --IW 200 --T 86 > gpurun_out/v3_microbench_icml.txt 2>&1
timeout 120 python tools/microbench.py --B 1024 --IW 128 --T 500 > gpurun_out/v3_microbench_large.txt 2>&1
timeout 120 python tools/step_timeline.py > gpurun_out/v3_step_timeline.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/v3_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline > gpurun_out/v3_launches.log 2>&1
timeout 250 bash tools/gpu_ncu3.sh v3_bwd_icml elbo_bwd 4
timeout 250 bash tools/gpu_ncu3.sh v3_fwd_icml elbo_fwd 4
timeout 250 bash tools/gpu_ncu.sh v3_bwd_large elbo_bwd --B 1024 --IW 128 --T 500
timeout 250 bash tools/gpu_ncu.sh v3_fwd_large elbo_fwd --B 1024 --IW 128 --T 500
rm -f gpurun_out/*_details.csv
ls -la gpurun_out | tail -30
