#!/bin/bash
set -u
mkdir -p gpurun_out
bash scripts/ab_bench.sh "" "DRN_CARVEOUT=1" "DRN_CARVEOUT=1 DRN_PDL=1" 2>&1 | tee gpurun_out/r02_ab_carveout.log
