#!/bin/bash
# vectorised pack / unpack of the k = 3 conv weights: full GPU suite, then A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu_final2.log
tail -2 gpurun_out/r02_pytest_gpu_final2.log
bash scripts/ab_bench.sh "" "DRN_PACK_V2=0" 2>&1 | tee gpurun_out/r02_ab_pack_v2.log
