#!/bin/bash
# one GPU-box call: parity tests, bench, eig timings, variant A/B, ncu launch list + full capture of the Chebyshev kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_r02.log
python bench.py > gpurun_out/bench_r02.log 2>gpurun_out/bench_r02.err; echo "bench rc=$?"
python tools/eig_bench.py > gpurun_out/eig_bench_r02.log 2>&1; echo "eig rc=$?"; cat gpurun_out/eig_bench_r02.log
for v in "" mb4; do
  if [ -n "$v" ]; then export TMQ_LIB_PATH=$PWD/quda-qkxtm-multigrid-plugin_b200/lib/variants/libtmq_$v.so; fi
  echo "variant=[$v]"; python tools/sweep.py --tiles "4,4,2" --precs 8,4 --recons 12 --reps 30
done > gpurun_out/variant_ab_r02.log 2>&1
unset TMQ_LIB_PATH
cat gpurun_out/variant_ab_r02.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python tools/eig_bench.py --skip-eig > gpurun_out/ncu_list_r02.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dslash_kernel -s 40 -c 6 -o gpurun_out/prof_r02 -f python tools/eig_bench.py --skip-eig > gpurun_out/ncu_full_r02.log 2>&1; echo "ncu full rc=$?"
