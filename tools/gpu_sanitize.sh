#!/bin/bash
# compute-sanitizer over every kernel family (tools/sanitize_driver.py): racecheck where kernels hand data between warps
# through shared memory (named-barrier rings, team prologue), memcheck everywhere.  Logs -> gpurun_out/ (summaries are
# copied to profiles/).  usage: gpu_sanitize.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
run() {
  tool=$1; which=$2
  log=gpurun_out/sanitize_${tag}_${tool}_${which}.log
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py $which > $log 2>&1
  echo "$tool $which: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_DRIVER_DONE' $log | tr '\n' ' ')"
}
for which in dr_latency dr_ws relay_precisions blackbox_mma step; do run racecheck $which; done
for which in dr_latency dr_ws dr_throughput relay_precisions hidden_precisions blackbox_mma blackbox_scalar exchange step; do run memcheck $which; done
