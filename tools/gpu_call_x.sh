#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_contract.py tests/test_gpu_host_shim.py -q > gpurun_out/pytest_gpu_r22.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu_r22.log
