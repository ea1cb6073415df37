#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29517 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 1 2 > gpurun_out/shard_t2_r05.log 2>&1; echo "shard T rc=$?"; grep halo_mode gpurun_out/shard_t2_r05.log
$TR --master-port 29518 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 2 1 --eig 0 > gpurun_out/shard_z2_r05.log 2>&1; echo "shard Z rc=$?"; grep halo_mode gpurun_out/shard_z2_r05.log
$TR --master-port 29519 bench.py --gpus 2 --steps 30 --warmup 3 --no-e2e > gpurun_out/bench48_n2_r05.log 2>&1; echo "bench n2 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench48_n2_r05.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['config']['halo'], {k:v['ms'] for k,v in d['kernels'].items()})
PY
