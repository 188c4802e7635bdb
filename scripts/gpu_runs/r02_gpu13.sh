#!/bin/bash
set -u
mkdir -p gpurun_out
bash scripts/ab_bench.sh "" "DRN_GRAD_VIEWS=0" 2>&1 | tee gpurun_out/r02_ab_hostpath.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_multistep_gpu.py -q -x 2>&1 | tail -3
