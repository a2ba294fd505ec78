#!/bin/bash
# round 2, pass L: GEMM path of the encoder at large batches: tests, bench, launch list of the synthetic workload
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 100 --warmup 5 2>gpurun_out/r02l_bench.err | tail -1 > gpurun_out/r02l_bench_1gpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l_bench_1gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernels')}); print(d['e2e'])
for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), round(v['value']/1e6,2), v['kernels'])
PY
bash tools/gpu_r2_k.sh 2>&1 | tail -14
