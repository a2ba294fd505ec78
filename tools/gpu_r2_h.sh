#!/bin/bash
# round 2, pass H: fused IWAE / fused lin-wgrad+Adam: GPU test-suite, bench, launch list
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 100 --warmup 5 2>gpurun_out/r02h_bench.err | tail -1 > gpurun_out/r02h_bench_1gpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_bench_1gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernels','gpu_launches','fusions')}); print(d['e2e'])
for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), v['kernels'], v.get('fusions'))
PY
VIHDS_FUSE_IWAE=0 VIHDS_FUSE_ADAM=0 timeout 300 python bench.py --steps 100 --warmup 5 --no-extra-workloads --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('unfused', d['ms_per_step'], d['e2e']['ms_per_step'], d['fusions'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02h_launches_bench.csv \
  python bench.py --steps 5 --warmup 3 --spin 0 --no-cpu-baseline --no-extra-workloads > gpurun_out/r02h_launches.log 2>&1
tail -8 gpurun_out/r02h_launches_bench.csv | cut -d, -f5,15
