#!/bin/bash
# round 2, pass I (2 GPUs): GPU test-suite incl. the 2-rank exchange test, bench at N=2
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r02i_bench_2gpu.json 2> gpurun_out/r02i_bench_2gpu.err; tail -2 gpurun_out/r02i_bench_2gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02i_bench_2gpu.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','params_identical_across_ranks','exchange_timed_out','fusions','gpu_launches')})
for k,v in d.get('workloads',{}).items(): print(k, v['ms_per_step'], v['value'], v.get('params_identical_across_ranks'), v.get('fusions'))
PY
