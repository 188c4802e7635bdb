#!/bin/bash
# Round bench + profiling pass (run under gpurun from the repo root).  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
kill $SMI
tail -c 3000 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
tail -c 1000 gpurun_out/bench_ref.json
# launch list (cold-cache, serialised: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
# full capture of the dominant kernel: gemm_pair launch #5 = prop_fc forward of the 2nd warm-up step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 4 -c 1 -o gpurun_out/prof_propfc \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
