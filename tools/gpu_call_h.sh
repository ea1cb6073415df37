#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29517 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 1 2 --eig 0 --pack-async 1 > gpurun_out/shard_t2_pa.log 2>&1; echo "shard T pack-async rc=$?"
$TR --master-port 29518 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 2 1 --eig 0 --pack-async 1 > gpurun_out/shard_z2_pa.log 2>&1; echo "shard Z pack-async rc=$?"
for pa in 0 1; do
for lat in "48 48 48 96" "48 48 48 24"; do
$TR --master-port 29519 bench.py --gpus 2 --steps 40 --warmup 3 --no-e2e --lattice $lat --pack-async $pa > gpurun_out/bench_n2_pa.log 2>&1; echo "bench rc=$?"
python - "$pa" "$lat" <<'PY'
import json, sys
for l in open('gpurun_out/bench_n2_pa.log'):
    if l.startswith('{'):
        d=json.loads(l); print('pack_async', sys.argv[1], 'lattice', sys.argv[2], 'ms/step', round(d['ms_per_step'],4), {k:round(v['ms'],4) for k,v in d['kernels'].items()})
PY
done; done
