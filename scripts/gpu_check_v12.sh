#!/bin/bash
# r01 v12: parity suite, bench, launch list, diagnostic ncu sections of the small kernels, configs[2]/[4] measurements
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu_v12.log
tail -4 gpurun_out/pytest_gpu_v12.log
bash scripts/ab_bench.sh "" 2>&1 | tee gpurun_out/ab_v12.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_v12.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_v12.log 2>&1
echo "ncu launches rc=$?"
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section Occupancy --section MemoryWorkloadAnalysis --section LaunchStats \
    --clock-control none -k regex:'qe_attn_bwd_scalars|col_stats|bn_bwd_apply|bn_relu_apply|lstm_fwd|sgemm_multi|head_proj' --launch-skip 60 -c 40 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_small_v12.txt 2>&1
echo "ncu small rc=$?"
timeout 600 python scripts/configs_bench.py --steps 10 --warmup 4 > gpurun_out/configs_v12.jsonl 2> gpurun_out/configs_v12.err
echo "configs rc=$?"; tail -3 gpurun_out/configs_v12.err; cat gpurun_out/configs_v12.jsonl | cut -c1-300
