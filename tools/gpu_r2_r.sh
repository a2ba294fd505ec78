#!/bin/bash
# quick: bench (primary workload only) + ncu of the mx kernel
timeout 600 python bench.py --steps 100 --warmup 5 --no-extra-workloads --no-cpu-baseline 2>gpurun_out/r02r_bench.err | tail -1 > gpurun_out/r02r_bench_1gpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02r_bench_1gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernels')}); print({k:d['e2e'][k] for k in ('value','ms_per_step','latency_ms')})
PY
B="python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads"
timeout 300 bash tools/gpu_ncu_cmd.sh r02r_mx_bwd_icml elbo_bwd_mx 6 $B > /dev/null
rm -f gpurun_out/*_details.csv
python tools/ncu_brief.py gpurun_out/ncu_r02r_mx_bwd_icml_raw.csv 2>/dev/null | head -8
