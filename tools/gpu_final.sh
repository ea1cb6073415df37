#!/bin/bash
# One GPU: the whole -m gpu suite, smoke(), both bench arms, then the ncu launch list and one full capture of the Dslash kernels.
# usage (under gpurun): bash tools/gpu_final.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out
export TMQ_HALO_TIMEOUT_MS=10000
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench48.json 2> $OUT/${TAG}_bench48.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench48_reference.json 2> $OUT/${TAG}_bench48_reference.err; echo "reference rc=$?"; tail -c 600 $OUT/${TAG}_bench48_reference.json
timeout 600 bash tools/gpu_profile.sh $TAG --scale64 0
python - <<PY
import json
b=json.loads([l for l in open('$OUT/${TAG}_bench48.json') if l.startswith('{')][-1])
print('value',b['value'],'ms',b['ms_per_step'],'solver',b['solver_loop']['ms_per_iter'],'roofline',b['roofline']['frac'],'e2e',b['e2e']['value'],b['e2e'].get('host_link'),'cpu',b['cpu_baseline']['value'])
print('scale64',b['scale64']['ms_per_step'],b['scale64']['solution_checksum'])
PY
