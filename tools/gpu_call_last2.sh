#!/bin/bash
mkdir -p gpurun_out
timeout 22 python -m pytest tests/test_gpu_contract.py -k "fixture" -x -q > gpurun_out/pytest_gpu_r28_passes.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_r28_passes.log | cut -c1-300
