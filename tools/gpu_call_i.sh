#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_clover.py -x -q > gpurun_out/pytest_gpu_r08.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu_r08.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29517 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 1 2 --eig 0 > gpurun_out/shard_t2_r08.log 2>&1; echo "shard T rc=$?"; grep -o "rank [0-9]/2[^;]*;[^;]*; failures: \[[^]]*\]" gpurun_out/shard_t2_r08.log
$TR --master-port 29518 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 2 1 --eig 0 > gpurun_out/shard_z2_r08.log 2>&1; echo "shard Z rc=$?"; grep -o "rank [0-9]/2[^;]*;[^;]*; failures: \[[^]]*\]" gpurun_out/shard_z2_r08.log
tail -5 gpurun_out/shard_t2_r08.log | cut -c1-600
