#!/bin/bash
# Multi-GPU parity on N = 2 GPUs (run under:  gpurun --gpus 2 -- 'bash tools/gpu_sharded.sh r2'):
#   tests/sharded_parity.py  every rank compares its slab of hop / M^dag M / CG / eigensolver / clover with the CPU oracle on the global
#                            lattice, T split and Z split, in the fused halo mode (3), the copy-engine mode (2) and over NCCL (0)
#   tests/sharded_shim.py    the C++ QKXTM shim under torchrun --no-python: invertQuda slabs, two- / three-point files, plaquette through
#                            the containers' ghost zones
TAG=${1:-r2}
OUT=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29517
for p2p in 3 2 0; do
  for grid in "1 1 1 2" "1 1 2 1"; do
    g=$(echo $grid | tr -d ' ')
    timeout 600 $TR --master-port $port tests/sharded_parity.py --lattice 8 8 8 16 --grid $grid --p2p $p2p --eig $([ $p2p = 3 ] && echo 1 || echo 0) > $OUT/shard_${TAG}_p${p2p}_g$g.log 2>&1
    echo "sharded_parity p2p=$p2p grid=$grid rc=$?"; grep -o "rank [0-9]/2[^;]*;[^;]*; failures: \[[^]]*\]" $OUT/shard_${TAG}_p${p2p}_g$g.log | cut -c1-300
    port=$((port+1))
  done
done
timeout 900 python tests/sharded_shim.py > $OUT/sharded_shim_$TAG.log 2>&1; echo "sharded shim rc=$?"; tail -8 $OUT/sharded_shim_$TAG.log | cut -c1-400
