#!/bin/bash
# final validation of the round: full GPU suite, smoke, bench (as the driver runs it), ncu launch list + full capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r20.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r20.log
python __graft_entry__.py smoke > gpurun_out/smoke_r20.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r20.log
python bench.py > gpurun_out/bench_r20.log 2> gpurun_out/bench_r20.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_r20.log
bash tools/gpu_profile.sh r20
