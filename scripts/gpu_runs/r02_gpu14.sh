#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/launch_gap_probe.py 2>&1 | tail -1
DRN_PDL=1 python scripts/launch_gap_probe.py 2>&1 | tail -1
bash scripts/ab_bench.sh "" "DRN_PDL=1" 2>&1 | tee gpurun_out/r02_ab_pdl.log
timeout 1200 python scripts/r1_parity_seeds.py --seeds 5 --steps 120 --eval-batches 32 --arms cuda > gpurun_out/r1_cuda_arm_120.json 2> gpurun_out/r02_r1_parity_120.err
tail -5 gpurun_out/r02_r1_parity_120.err; python -c "
import json; d=json.load(open('gpurun_out/r1_cuda_arm_120.json')); print(json.dumps(d['R@1']))"
