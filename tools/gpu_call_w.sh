#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_contract.py -q > gpurun_out/pytest_gpu_r21.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu_r21.log
