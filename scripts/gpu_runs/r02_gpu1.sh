#!/bin/bash
# r02 first GPU pass: stream-K parity + A/B against the static schedule
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest_gemm.log
cat gpurun_out/r02_pytest_gemm.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_a.log
cat gpurun_out/r02_pytest_gpu_a.log
bash scripts/ab_bench.sh "DRN_STREAMK=0" "DRN_STREAMK=1" "DRN_STREAMK=1 DRN_SK_MIN=8" 2>&1 | tee gpurun_out/r02_ab_streamk.log
DRN_STREAMK=0 timeout 300 python scripts/insitu_timeline.py > gpurun_out/r02_insitu_static.json 2> gpurun_out/r02_insitu_static.err
DRN_STREAMK=1 timeout 300 python scripts/insitu_timeline.py > gpurun_out/r02_insitu_streamk.json 2> gpurun_out/r02_insitu_streamk.err
ls -la gpurun_out | tail -5
