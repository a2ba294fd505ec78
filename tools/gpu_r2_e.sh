#!/bin/bash
# round 2, pass E (2 GPUs): exchange kernel checks, then the bench at N=2 with all workloads
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 tools/peer_check.py 2>&1 | grep -v Warning | tail -24
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err; tail -3 gpurun_out/bench_r02_2gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02_2gpu.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','params_identical_across_ranks','exchange_timed_out','exchange','skipped_steps_nan_guard')})
for k,v in d.get('workloads',{}).items(): print(k, v['ms_per_step'], v['value'], v.get('params_identical_across_ranks'))
PY
