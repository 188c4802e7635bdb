#!/bin/bash
# r01 v14: programmatic dependent launch -- parity suite with it on, A/B against plain stream order
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu_v14.log
tail -4 gpurun_out/pytest_gpu_v14.log
bash scripts/ab_bench.sh "DRN_PDL=0" "DRN_PDL=1" 2>&1 | tee gpurun_out/ab_v14.log
