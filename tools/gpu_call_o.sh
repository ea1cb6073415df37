#!/bin/bash
mkdir -p gpurun_out
for v in "" s7 s8; do
  if [ -n "$v" ]; then export TMQ_LIB_PATH=$PWD/quda-qkxtm-multigrid-plugin_b200/lib/variants/libtmq_$v.so; else unset TMQ_LIB_PATH; fi
  echo "variant=[$v]"; python tools/sweep.py --tiles "4,4,2" --precs 4 --recons 12,18 --reps 40
done > gpurun_out/variant_fp32_r14.log 2>&1
cat gpurun_out/variant_fp32_r14.log
