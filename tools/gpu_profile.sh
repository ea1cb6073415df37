#!/bin/bash
# Run on the GPU box (under gpurun): plain bench, then the ncu launch list and one full capture of the Dslash kernel.
# usage: tools/gpu_profile.sh <tag> [extra bench args]
set -u
TAG=${1:-r01}; shift || true
OUT=gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu $*"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_list_$TAG.log 2>&1
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dslash_kernel -s 8 -c 4 -f -o $OUT/prof_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/ncu_full_$TAG.log
