#!/bin/bash
# full validation + profile evidence with the final kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r10.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_r10.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r10.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r10.log
python bench.py > gpurun_out/bench_r10.log 2>gpurun_out/bench_r10.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r10.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernels'].items()}, d['e2e']['value'], d['cpu_baseline']['value'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r10.csv python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_list_r10.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:dslash_kernel -s 30 -c 8 -o gpurun_out/prof_r10 -f python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_r10.log 2>&1; echo "ncu full rc=$?"
