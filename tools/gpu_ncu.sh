#!/bin/bash
# full ncu capture of one kernel of the microbench; exports the raw + source pages as (gzipped) csv and drops the .ncu-rep
# usage: gpu_ncu.sh <tag> <kernel-regex> <microbench args...>
set -x
tag=$1; shift
kre=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$kre -s 1 -c 1 -o /tmp/prof_$tag python tools/microbench.py "$@" --iters 1 > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page details --csv > gpurun_out/ncu_${tag}_details.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_${tag}_source.csv.gz
ls -la gpurun_out /tmp/prof_$tag.ncu-rep
