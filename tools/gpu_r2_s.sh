#!/bin/bash
# all-lane mbarrier arrival: racecheck of the matrix-form kernel + bench
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python tools/sanitize_driver.py dr_latency > gpurun_out/sanitize_r02s_racecheck_dr_latency.log 2>&1
grep -E 'RACECHECK SUMMARY|SANITIZE_DRIVER_DONE' gpurun_out/sanitize_r02s_racecheck_dr_latency.log
head -8 gpurun_out/sanitize_r02s_racecheck_dr_latency.log | cut -c1-220
timeout 600 python bench.py --steps 100 --warmup 5 --no-extra-workloads --no-cpu-baseline 2>gpurun_out/r02s_bench.err | tail -1 > gpurun_out/r02s_bench_1gpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s_bench_1gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernels')}); print({k:d['e2e'][k] for k in ('value','ms_per_step','latency_ms')})
PY
