#!/bin/bash
# r02 final single-GPU pass: full GPU test suite, smoke, bench (both arms), launch list, ncu captures
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu_final.log
tail -2 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r02_bench_final.json | head -c 1500
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref_final.json 2>> gpurun_out/r02_bench_final.err
tail -c 400 gpurun_out/r02_bench_ref_final.json
TAG=r02_final
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --sustain-seconds 0 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -c 18 -o /tmp/prof_pair_$TAG \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra --sustain-seconds 0 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i /tmp/prof_pair_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_pair_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_pair_$TAG.ncu-rep --page details > gpurun_out/ncu_pair_${TAG}_details.txt 2>/dev/null
bash scripts/gpu_ncu_small.sh $TAG
