#!/bin/bash
# 2 GPUs: parity of the fused-pack halo mode (4) and of the fused peer-store mode (3), T and Z split, then a quick 48^3x96 bench per mode
OUT=gpurun_out
export TMQ_HALO_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
port=29517
for p2p in 4 3; do
  for grid in "1 1 1 2" "1 1 2 1"; do
    g=$(echo $grid | tr -d ' ')
    timeout 300 $TR --master-port $port tests/sharded_parity.py --lattice 8 8 8 16 --grid $grid --p2p $p2p --eig 0 > $OUT/n2_p${p2p}_g$g.log 2>&1
    echo "p2p=$p2p grid=$grid rc=$?"; grep -o "failures: \[[^]]*\]" $OUT/n2_p${p2p}_g$g.log
    port=$((port+1))
  done
done
for halo in fusedce p2p fused; do
  timeout 400 $TR --master-port 29529 bench.py --gpus 2 --steps 20 --warmup 5 --halo $halo --no-cpu --no-e2e --scale64 $([ $halo = fused ] && echo 0 || echo 1) > $OUT/n2_bench_$halo.json 2> $OUT/n2_bench_$halo.err
  echo "bench $halo rc=$?"; python -c "
import json,sys
b=json.loads([l for l in open('$OUT/n2_bench_$halo.json') if l.startswith('{')][-1])
s=b.get('scale64') or {}
print('$halo', 'ms_per_step', b['ms_per_step'], 'value', b['value'], 'solver_loop', b['solver_loop']['ms_per_iter'], b['run']['halo'], '| s64', s.get('ms_per_step'), (s.get('solution_checksum') or {}).get('sum_x'), (s.get('solution_checksum') or {}).get('norm2_x'))"
done
