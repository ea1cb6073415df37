#!/bin/bash
# 2 GPUs: fused halo mode parity (T and Z split) and a quick 48^3x96 bench in the fused and the copy-engine mode
OUT=gpurun_out
export TMQ_HALO_TIMEOUT_MS=10000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29517 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 1 2 --p2p 3 --eig 0 > $OUT/n2_fused_t.log 2>&1; echo "T rc=$?"; grep -o "failures: \[[^]]*\]" $OUT/n2_fused_t.log
timeout 300 $TR --master-port 29518 tests/sharded_parity.py --lattice 8 8 8 16 --grid 1 1 2 1 --p2p 3 --eig 0 > $OUT/n2_fused_z.log 2>&1; echo "Z rc=$?"; grep -o "failures: \[[^]]*\]" $OUT/n2_fused_z.log
for halo in fused p2p; do
  timeout 400 $TR --master-port 29519 bench.py --gpus 2 --steps 20 --warmup 5 --halo $halo --no-cpu --no-e2e --scale64 0 > $OUT/n2_bench_$halo.json 2> $OUT/n2_bench_$halo.err
  echo "bench $halo rc=$?"; python -c "
import json,sys
b=json.load(open('$OUT/n2_bench_$halo.json'))
print('$halo', 'ms_per_step', b['ms_per_step'], 'value', b['value'], 'solver_loop', b['solver_loop']['ms_per_iter'], b['run']['halo'])"
done
