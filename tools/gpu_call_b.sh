#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_smear.py tests/test_gpu_host_shim.py tests/test_gpu_eig.py -x -q > gpurun_out/pytest_gpu_r03.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu_r03.log
python tools/smear_bench.py > gpurun_out/smear_bench_r03.log 2>&1; echo "smear rc=$?"; cat gpurun_out/smear_bench_r03.log
python bench.py --no-cpu > gpurun_out/bench_r03.log 2>gpurun_out/bench_r03.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r03.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['frac'], {k:v['ms'] for k,v in d['kernels'].items()}, d['e2e']['value'])
PY
ncu --set full --clock-control none --import-source on -k regex:gauss_smear -s 4 -c 3 -o gpurun_out/prof_r03_smear -f python tools/smear_bench.py --nsmear 4 --blocks 1000000 --precs 8 > gpurun_out/ncu_full_r03.log 2>&1; echo "ncu full rc=$?"
