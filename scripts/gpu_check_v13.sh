#!/bin/bash
# r01 v13: parity suite, bench, launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_v13.log
tail -4 gpurun_out/pytest_gpu_v13.log
bash scripts/ab_bench.sh "" 2>&1 | tee gpurun_out/ab_v13.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_v13.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_v13.log 2>&1
echo "ncu launches rc=$?"
