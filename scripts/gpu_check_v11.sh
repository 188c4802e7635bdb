#!/bin/bash
# r01 v11: parity suite, A/B of the side-branch schedule, launch list (run under gpurun from the repo root)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_v11.log
tail -8 gpurun_out/pytest_gpu_v11.log
bash scripts/ab_bench.sh "DRN_SIDE=0" "DRN_SIDE=1" 2>&1 | tee gpurun_out/ab_v11.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_v11.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_v11.log 2>&1
echo "ncu launches rc=$?"
