#!/bin/bash
# r01 v15: BatchNorm statistics fused into the contraction epilogue -- parity suite, A/B, launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu_v15.log
tail -4 gpurun_out/pytest_gpu_v15.log
bash scripts/ab_bench.sh "DRN_FUSED_STATS=0" "DRN_FUSED_STATS=1" 2>&1 | tee gpurun_out/ab_v15.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_v15.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_v15.log 2>&1
echo "ncu launches rc=$?"
