#!/bin/bash
# round 2, pass Q: tests, bench (all workloads), ncu of one kernel (default: the team forward kernel)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 100 --warmup 5 2>gpurun_out/r02q_bench.err | tail -1 > gpurun_out/r02q_bench_1gpu.json
tail -3 gpurun_out/r02q_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02q_bench_1gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernels')}); print({k:d['e2e'][k] for k in ('value','ms_per_step','latency_ms')})
for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['value']/1e6,2), v['kernels'], v['cost_after_last_step'], v['skipped_steps_nan_guard'])
PY
B="python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads"
timeout 300 bash tools/gpu_ncu_cmd.sh r02q_team_fwd_icml elbo_fwd_team 6 $B
rm -f gpurun_out/*_details.csv
python tools/ncu_brief.py gpurun_out/ncu_r02q_team_fwd_icml_raw.csv 2>/dev/null | head -12
