#!/bin/bash
# full ncu capture of one kernel of a bench.py run (tag, kernel regex, launch-skip), csv export on the box
tag=$1; kre=$2; skip=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -o /tmp/prof_$tag python bench.py --steps 2 --warmup 1 --spin 0 --no-cpu-baseline "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_${tag}_source.csv.gz
