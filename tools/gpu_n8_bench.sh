#!/bin/bash
# 8 GPUs: T x Z parity of the default halo mode, then 48^3x96 strong scaling + e2e + the 64^3x128 T x Z leg in the default (fused pack +
# copy-engine push) and the copy-engine mode
OUT=gpurun_out
TAG=${1:-r2}
export TMQ_HALO_TIMEOUT_MS=20000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29518 tests/sharded_parity.py --lattice 8 8 16 32 --grid 1 1 2 4 --p2p 4 --eig 0 > $OUT/n8_${TAG}_parity_z2t4.log 2>&1; echo "parity TxZ rc=$?"
grep -o "rank [0-9]/8[^;]*;[^;]*; failures: \[[^]]*\]" $OUT/n8_${TAG}_parity_z2t4.log | cut -c1-200
timeout 600 $TR --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 --halo fusedce --no-cpu > $OUT/n8_${TAG}_fusedce.json 2> $OUT/n8_${TAG}_fusedce.err; echo "fusedce rc=$?"
timeout 500 $TR --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 --halo p2p --no-cpu --no-e2e --scale64 0 > $OUT/n8_${TAG}_p2p.json 2> $OUT/n8_${TAG}_p2p.err; echo "p2p rc=$?"
python - <<PY
import json
for h in ('fusedce','p2p'):
    try:
        txt=open('$OUT/n8_${TAG}_%s.json'%h).read()
        b=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
        s=b.get('scale64') or {}
        e=b.get('e2e') or {}
        print(h,'ms',round(b['ms_per_step'],4),'value',round(b['value']),'solver',round(b['solver_loop']['ms_per_iter'],4),'K2',round(b['kernels']['K2 hop+A^-1']['ms'],4),
              '| s64 ms',s.get('ms_per_step'),'sum',(s.get('solution_checksum') or {}).get('sum_x'),(s.get('solution_checksum') or {}).get('norm2_x'),'| e2e',e.get('value'),(e.get('single_solve') or {}).get('value'))
    except Exception as ex:
        print(h,'parse failed',ex)
PY
tail -3 $OUT/n8_${TAG}_fusedce.err | cut -c1-300
