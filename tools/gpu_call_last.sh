#!/bin/bash
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu_r27_parity.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r27_parity.log | cut -c1-200
