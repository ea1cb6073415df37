#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_contract.py tests/test_gpu_smear.py -x -q > gpurun_out/pytest_gpu_r16.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r16.log
python tools/contract_bench.py > gpurun_out/contract_bench_r16.log 2>&1; echo "bench rc=$?"; cut -c100-330 gpurun_out/contract_bench_r16.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_contract_r16.csv python tools/contract_bench.py --precs 4 --qsq 16 > gpurun_out/ncu_contract_r16.log 2>&1
grep -E "meson_site|axis_dft" gpurun_out/launches_contract_r16.csv | cut -d, -f5,12- | tail -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_contract8_r16.csv python tools/contract_bench.py --precs 8 --qsq 16 > gpurun_out/ncu_contract8_r16.log 2>&1
grep -E "meson_site|axis_dft" gpurun_out/launches_contract8_r16.csv | cut -d, -f5,12- | tail -8
