#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x 2>&1 | tail -2
bash scripts/ab_bench.sh "" "DRN_HYBRID_MIN_SAVED=8" 2>&1 | tee gpurun_out/r02_ab_hybrid_threshold.log
