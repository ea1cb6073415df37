#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_eig.py tests/test_gpu_host_shim.py -x -q > gpurun_out/pytest_gpu_r09.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_r09.log
