#!/bin/bash
# scaling run on N GPUs of one box (N = number visible): bench.py under torchrun with all workloads
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; tail -2 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_${N}gpu.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','params_identical_across_ranks','exchange_timed_out','fusions')})
for k,v in d.get('workloads',{}).items(): print(k, v['ms_per_step'], v['value'], v.get('params_identical_across_ranks'))
PY
