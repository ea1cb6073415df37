#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -k "reconstruct_8" -q > gpurun_out/pytest_gpu_r26_recon8.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu_r25_recon8.log | cut -c1-300
timeout 60 python tools/sweep.py --tiles "4,4,2" --precs 8,4 --recons 8,12 --reps 30 > gpurun_out/sweep_recon8_r26.log 2>&1; echo "sweep rc=$?"; cat gpurun_out/sweep_recon8_r26.log | cut -c1-300
