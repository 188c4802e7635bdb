#!/bin/bash
# Round bench + profiling pass (run under gpurun from the repo root).  Outputs land in gpurun_out/ (kept small: the
# .ncu-rep stays on the box, only CSV / text exports come back).  Usage: bash scripts/gpu_bench_profile.sh <tag>
set -u
TAG=${1:-v}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_$TAG.log
tail -2 gpurun_out/pytest_gpu_$TAG.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$TAG.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"
kill $SMI
tail -c 3000 gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
tail -c 600 gpurun_out/bench_ref_$TAG.json
# launch list (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
# full capture of the first 18 pair-kernel launches (= the first, eager, step: 16 per step; prop_fc forward is the first
# 148-CTA launch of ~0.5 ms, prop_fc weight gradient the last one)
timeout 900 ncu --set full --clock-control none -k regex:gemm_pair_kernel -c 18 -o /tmp/prof_pair_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/prof_pair_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_pair_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_pair_$TAG.ncu-rep --page details > gpurun_out/ncu_pair_${TAG}_details.txt 2>/dev/null
ls -la gpurun_out | tail -8
