#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_contract.py tests/test_gpu_smear.py -x -q > gpurun_out/pytest_gpu_r15.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_r15.log
python tools/contract_bench.py > gpurun_out/contract_bench_r15.log 2>&1; echo "bench rc=$?"; cat gpurun_out/contract_bench_r15.log | cut -c1-400
