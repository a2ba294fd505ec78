#!/bin/bash
# full ncu capture of one kernel of an arbitrary command; exports the raw / details / source pages as csv (source gzipped)
# usage: gpu_ncu_cmd.sh <tag> <kernel-regex> <skip> <command...>
tag=$1; shift
kre=$1; shift
skip=$1; shift
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -o /tmp/prof_$tag "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page details --csv > gpurun_out/ncu_${tag}_details.csv 2>/dev/null
ncu -i /tmp/prof_$tag.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_${tag}_source.csv.gz
ls -la gpurun_out/ncu_${tag}* /tmp/prof_$tag.ncu-rep
