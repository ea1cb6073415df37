#!/bin/bash
# 8 GPUs, final build: the bench line as the driver runs it (default halo mode, e2e with the host-link probe, 64^3x128 T x Z leg), then
# the e2e leg again with the CG's host loop three iterations ahead of the residual read-back
OUT=gpurun_out
TAG=${1:-r2c}
N=${2:-8}
export TMQ_HALO_TIMEOUT_MS=20000
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > $OUT/n${N}_${TAG}.json 2> $OUT/n${N}_${TAG}.err; echo "bench rc=$?"
TMQ_CG_LAG=3 timeout 400 $TR --master-port 29520 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --scale64 0 > $OUT/n${N}_${TAG}_lag3.json 2> $OUT/n${N}_${TAG}_lag3.err; echo "lag3 rc=$?"
python - <<PY
import json
for h in ('','_lag3'):
    try:
        txt=open('$OUT/n${N}_${TAG}%s.json'%h).read()
        b=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
        s=b.get('scale64') or {}
        e=b.get('e2e') or {}
        print(h or 'default','ms',round(b['ms_per_step'],4),'value',round(b['value']),'solver',round(b['solver_loop']['ms_per_iter'],4),
              '| s64 ms',s.get('ms_per_step'),'sum',(s.get('solution_checksum') or {}).get('sum_x'),'| e2e',e.get('value'),'secs',e.get('secs'),'solver_secs',e.get('solver_secs'),
              'single',(e.get('single_solve') or {}).get('secs'),'mixed',(e.get('single_solve_mixed') or {}).get('secs'),'link',{k:v for k,v in (e.get('host_link') or {}).items() if k!='what'})
    except Exception as ex:
        print(h,'parse failed',ex)
PY
tail -2 $OUT/n${N}_${TAG}.err | cut -c1-300
