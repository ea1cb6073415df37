#!/bin/bash
# 8-GPU box: strong scaling of the fused CG iteration, copy-engine halo path (default), 48^3x96 and 64^3x128
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # name nproc args...
  local name=$1 np=$2; shift 2
  $TR --nproc-per-node $np --master-port $((29600 + RANDOM % 300)) bench.py --gpus $np --steps 40 --warmup 3 --no-e2e "$@" > gpurun_out/$name.log 2>&1
  echo "$name rc=$?"
  python - "$name" <<'PY'
import json, sys
for l in open('gpurun_out/%s.log' % sys.argv[1]):
    if l.startswith('{'):
        d = json.loads(l); print(sys.argv[1], d['config']['workload'][:14], d['config']['grid'], d['config']['halo'][:12], 'ms/step', round(d['ms_per_step'], 4), 'GF', round(d['value']), {k: round(v['ms'], 4) for k, v in d['kernels'].items()})
PY
}
run bench48_n8_r07 8
run bench64_n8_r07 8 --lattice 64 64 64 128
run bench64_n8_z2t4_r07 8 --lattice 64 64 64 128 --grid 1 1 2 4
run bench64_n8_nccl_r07 8 --lattice 64 64 64 128 --halo nccl
run bench48_n4_r07 4
run bench64_n4_r07 4 --lattice 64 64 64 128
$TR --nproc-per-node 8 --master-port 29517 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 2 4 --eig 0 > gpurun_out/shard_n8_r07.log 2>&1; echo "shard n8 rc=$?"; grep -c "failures: \[\]" gpurun_out/shard_n8_r07.log
