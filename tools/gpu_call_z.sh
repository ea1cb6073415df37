#!/bin/bash
# end-of-round check at HEAD: the full GPU suite and smoke()
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r24.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu_r24.log
python __graft_entry__.py smoke > gpurun_out/smoke_r24.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r24.log
