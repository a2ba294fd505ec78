#!/bin/bash
# round 2, pass M: double-buffered input sets of the end-to-end step (streamed e2e), optimiser state restore in the bench,
# FFMA2 issue-rate microbenchmark
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 100 --warmup 5 2>gpurun_out/r02m_bench.err | tail -1 > gpurun_out/r02m_bench_1gpu.json
tail -3 gpurun_out/r02m_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02m_bench_1gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernels')}); print(d['e2e'])
for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['value']/1e6,2), v['kernels'], v['cost_after_last_step'], v['skipped_steps_nan_guard'])
PY
timeout 120 tools/micro/ffma2_rate 2>&1 | tee gpurun_out/r02_ffma2_rate.txt
