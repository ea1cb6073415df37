#!/bin/bash
mkdir -p gpurun_out
timeout 500 python tests/sharded_shim.py > gpurun_out/sharded_shim_r23.log 2>&1; echo "sharded shim rc=$?"; tail -25 gpurun_out/sharded_shim_r23.log | cut -c1-600
