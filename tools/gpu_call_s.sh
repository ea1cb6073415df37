#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_contract.py -x -q > gpurun_out/pytest_gpu_r17.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_r17.log
python tools/contract_bench.py --qsq 3,16 > gpurun_out/contract_bench_r17.log 2>&1; echo "bench rc=$?"; cut -c100-330 gpurun_out/contract_bench_r17.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8 --csv --log-file gpurun_out/launches_contract_r17.csv python tools/contract_bench.py --precs 4 --qsq 3 > gpurun_out/ncu_contract_r17.log 2>&1
grep -E "meson_site" gpurun_out/launches_contract_r17.csv | cut -d, -f12- | tail -3
