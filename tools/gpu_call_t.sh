#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_contract.py -x -q > gpurun_out/pytest_gpu_r18.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu_r18.log
timeout 240 python tools/contract_bench.py --lattice 32 32 32 64 --qsq 3 --baryons 1 > gpurun_out/contract_bench_r18.log 2>&1; echo "bench rc=$?"; cut -c1-120,180-420 gpurun_out/contract_bench_r18.log
