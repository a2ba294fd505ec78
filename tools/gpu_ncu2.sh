#!/bin/bash
# usage: gpu_ncu2.sh <tag> <kernel-regex> <python script + args...>   (full ncu capture of one launch, csv export on the box)
tag=$1; shift; kre=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$kre -s 1 -c 1 -o /tmp/prof_$tag python "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page details --csv > gpurun_out/ncu_${tag}_details.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_${tag}_source.csv.gz
