#!/bin/bash
# One GPU call that regenerates the r02 evidence: the bench line (all workloads), the launch list of bench.py, full ncu
# captures (csv-exported on the box) of the dr_blackbox tensor-core kernels at the icml size, and the scalar black-box
# kernels' instruction totals for the before / after table.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python bench.py --steps 100 --warmup 5 2>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 8 --warmup 2 2>>gpurun_out/r02_bench.err | tail -1 > gpurun_out/r02_bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads > gpurun_out/r02_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_blackbox.csv \
  python bench.py --workload dr_blackbox_icml --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads > gpurun_out/r02_launches_bb.log 2>&1
export VIHDS_BB_IMPL=mma
timeout 300 bash tools/gpu_ncu_cmd.sh r02_bbm_fwd_icml bbm_fwd 1 python tools/bb_microbench.py --B 36 --IW 200 --iters 1
timeout 300 bash tools/gpu_ncu_cmd.sh r02_bbm_bwd_icml bbm_bwd 1 python tools/bb_microbench.py --B 36 --IW 200 --iters 1
timeout 300 bash tools/gpu_ncu_cmd.sh r02_bbm_bwd_large bbm_bwd 1 python tools/bb_microbench.py --B 1024 --IW 128 --iters 1
export VIHDS_BB_IMPL=scalar
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:bb_ -c 4 --csv --log-file gpurun_out/r02_bb_scalar_insts.csv \
  python tools/bb_microbench.py --B 36 --IW 200 --iters 1 > /dev/null 2>&1
rm -f gpurun_out/*_details.csv
ls -la gpurun_out | tail -20
