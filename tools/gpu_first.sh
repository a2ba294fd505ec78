#!/bin/bash
# first GPU pass: parity tests, smoke, kernel microbench at icml size and at a large slab, launch list + full ncu of the reverse kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python tools/microbench.py --B 36 --IW 200 --T 86 2>&1 | tail -4
python tools/microbench.py --B 1024 --IW 128 --T 500 --iters 5 2>&1 | tail -4
python tools/microbench.py --B 8192 --IW 128 --T 86 --iters 5 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_microbench.csv python tools/microbench.py --B 1024 --IW 128 --T 500 --iters 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:elbo_bwd -s 1 -c 1 -o gpurun_out/prof_bwd_large python tools/microbench.py --B 1024 --IW 128 --T 500 --iters 1 > gpurun_out/ncu_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:elbo_fwd -s 1 -c 1 -o gpurun_out/prof_fwd_large python tools/microbench.py --B 1024 --IW 128 --T 500 --iters 1 > gpurun_out/ncu_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:elbo_bwd -s 1 -c 1 -o gpurun_out/prof_bwd_icml python tools/microbench.py --B 36 --IW 200 --T 86 --iters 1 > gpurun_out/ncu_bwd_icml.log 2>&1
ls -la gpurun_out
