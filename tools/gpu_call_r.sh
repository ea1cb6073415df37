#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29517 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 1 2 --eig 0 > gpurun_out/shard_t2_r16.log 2>&1; echo "shard T rc=$?"; grep -o "rank [0-9]/2[^;]*;[^;]*; failures: \[[^]]*\]" gpurun_out/shard_t2_r16.log
$TR --master-port 29518 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 2 1 --eig 0 > gpurun_out/shard_z2_r16.log 2>&1; echo "shard Z rc=$?"; grep -o "rank [0-9]/2[^;]*;[^;]*; failures: \[[^]]*\]" gpurun_out/shard_z2_r16.log
tail -4 gpurun_out/shard_z2_r16.log | cut -c1-800
