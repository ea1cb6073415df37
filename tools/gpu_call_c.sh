#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r04.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu_r04.log
python bench.py --no-cpu > gpurun_out/bench_r04.log 2>gpurun_out/bench_r04.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r04.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernels'].items()}, d['e2e'])
PY
python tools/sweep.py --tiles "4,4,2" --precs 8,4 --recons 12,18 --reps 30 > gpurun_out/sweep_r04.log 2>&1; cat gpurun_out/sweep_r04.log
python tools/smear_bench.py --blocks 1000000 > gpurun_out/smear_bench_r04.log 2>&1; cat gpurun_out/smear_bench_r04.log
python tools/eig_bench.py --skip-eig > gpurun_out/eig_bench_r04.log 2>&1; cat gpurun_out/eig_bench_r04.log
