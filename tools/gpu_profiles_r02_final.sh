#!/bin/bash
# One GPU call that regenerates the final round-2 evidence: GPU test suite, the bench line (all workloads) and the reference
# arm, launch lists (time + instructions per launch) of the icml / synthetic / black-box steps, full ncu captures of the two
# ODE kernels of the headline workload, compute-sanitizer over every kernel family.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r02f_pytest_gpu.txt
timeout 600 python bench.py --steps 100 --warmup 5 2>gpurun_out/r02f_bench.err | tail -1 > gpurun_out/r02f_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 8 --warmup 2 2>>gpurun_out/r02f_bench.err | tail -1 > gpurun_out/r02f_bench_reference_arm.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_1gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernels')}); print({k:d['e2e'][k] for k in ('value','ms_per_step','latency_ms')})
for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['value']/1e6,2), v['kernels'], v['cost_after_last_step'], v['skipped_steps_nan_guard'])
r=json.loads(open('gpurun_out/r02f_bench_reference_arm.json').read()); print('reference arm', r['value'], r['ms_per_step'])
PY
M="gpu__time_duration.sum,smsp__inst_executed.sum"
timeout 300 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads > gpurun_out/r02f_launches.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench_blackbox.csv \
  python bench.py --workload dr_blackbox_icml --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench_synthetic.csv \
  python bench.py --workload synthetic_dr_constant --steps 3 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads > /dev/null 2>&1
B="python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads"
timeout 300 bash tools/gpu_ncu_cmd.sh r02f_mx_bwd_icml elbo_bwd_mx 6 $B > /dev/null
timeout 300 bash tools/gpu_ncu_cmd.sh r02f_team_fwd_icml elbo_fwd_team 6 $B > /dev/null
timeout 300 bash tools/gpu_ncu_cmd.sh r02f_enc_fwd_icml enc_fwd_kernel 6 $B > /dev/null
timeout 300 bash tools/gpu_ncu_cmd.sh r02f_enc_bwd_icml enc_bwd_kernel 6 $B > /dev/null
rm -f gpurun_out/*_details.csv
bash tools/gpu_sanitize.sh r02f 2>&1 | tail -16
