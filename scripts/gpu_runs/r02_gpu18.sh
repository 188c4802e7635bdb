#!/bin/bash
# r02 ncu evidence of the current kernels (8 epilogue warps): launch list, full capture of the contraction launches, small-kernel sections
set -u
TAG=r02_epi8
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --sustain-seconds 0 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -c 18 -o /tmp/prof_pair_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --sustain-seconds 0 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/prof_pair_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_pair_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_pair_$TAG.ncu-rep --page details > gpurun_out/ncu_pair_${TAG}_details.txt 2>/dev/null
bash scripts/gpu_ncu_small.sh $TAG
ls -la gpurun_out | tail -8
