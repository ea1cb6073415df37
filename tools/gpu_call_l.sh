#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_clover.py tests/test_gpu_host_shim.py tests/test_gpu_smear.py -x -q > gpurun_out/pytest_gpu_r11.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r11.log
python tools/clover_bench.py > gpurun_out/clover_bench_r11.log 2>&1; echo "clover bench rc=$?"; cat gpurun_out/clover_bench_r11.log
