#!/bin/bash
# round 2, pass J: SFU sigmoid/tanh + kept NeuralPrecisions activations: GPU test-suite, bench
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 600 python bench.py --steps 100 --warmup 5 2>gpurun_out/r02j_bench.err | tail -1 > gpurun_out/r02j_bench_1gpu.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02j_bench_1gpu.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernels','gpu_launches')}); print(d['e2e'])
for k,v in d['workloads'].items(): print(k, round(v['ms_per_step'],4), v['kernels'])
PY
