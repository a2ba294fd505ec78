#!/bin/bash
# round 2, pass D: blackbox check + timing with 4-warp reverse CTAs; full bench at N=1
mkdir -p gpurun_out
timeout 600 python tools/bb_mma_check.py 2>&1 | tail -3
VIHDS_BB_IMPL=mma timeout 300 python tools/bb_microbench.py --B 36 --IW 200 2>&1 | tail -1
VIHDS_BB_IMPL=mma timeout 300 python tools/bb_microbench.py --B 1024 --IW 128 --iters 3 2>&1 | tail -1
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -3 gpurun_out/bench_r02_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r02_a.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernels','clocks')})
print(d['e2e']); print(d['roofline']['frac'], d['cpu_baseline'])
for k,v in d.get('workloads',{}).items(): print(k, v['ms_per_step'], v['value'], v['kernels'], v['roofline']['frac'], v['roofline_fwd']['frac'])
PY
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
